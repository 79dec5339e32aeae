import numpy as np, struct, random
f32=np.float32
def bits(x): return struct.unpack('<I',struct.pack('<f',float(x)))[0]
def frombits(b): return f32(struct.unpack('<f',struct.pack('<I',b&0xFFFFFFFF))[0])
def seq(s,d,T):
    k=0; last=None
    while s<T:
        last=s; s=f32(s+d); k+=1
    return s,k,last
def tie_eb(d):
    b=bits(d); m=(b&0x7FFFFF)|0x800000; tz=(m&-m).bit_length()-1
    e=(b>>23)&0xFF            # biased exponent of d; lowest set bit has biased exponent e-23+tz
    # tie at binade with biased exponent E iff E-24 == e-23+tz  -> E = e+1+tz
    return ((e+1+tz)&0xFF)<<23
def jump(s,d,T):
    k=0; last=None; it=0
    teb=tie_eb(d)
    while s<T:
        it+=1
        s1=f32(s+d); q=f32(s1-s)
        sb=bits(s); eb=sb&0x7F800000
        top=frombits(eb+0x00800000)
        if s1<top and (eb!=teb or (sb&1)==0):
            hi=min(T,top)
            est=f32(f32(hi-s)/q)*f32(0.99999)
            jf=f32(np.floor(est))
            sj=f32(np.float64(jf)*np.float64(q)+np.float64(s))
            n=0
            while f32(sj+q)<hi:
                sj=f32(sj+q); jf=f32(jf+1); n+=1
            assert n<=2
            last=sj; s=f32(sj+d); k+=int(jf)+1
        else:
            last=s; s=s1; k+=1
    return s,k,last,it
random.seed(2)
hist={}
for itn in range(300000):
    d=f32(10**random.uniform(-3,3))
    if random.random()<0.3:
        b=bits(d)&~((1<<random.randint(0,22))-1); d=frombits(b)
    s=f32(random.uniform(0,1)*float(d)) if random.random()<0.5 else f32(float(d)*random.uniform(0,3000))
    T=f32(float(s)+float(d)*random.uniform(0,5000)*random.random()**3)
    a=seq(s,d,T); b=jump(s,d,T)
    assert a[0]==b[0] and a[1]==b[1] and (a[2]==b[2]), (s,d,T,a,b)
    hist[b[3]]=hist.get(b[3],0)+1
print("ok",sorted(hist.items()))
