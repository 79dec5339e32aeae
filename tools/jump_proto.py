import numpy as np, struct, random
f32=np.float32
def bits(x): return struct.unpack('<I',struct.pack('<f',float(x)))[0]
def frombits(b): return f32(struct.unpack('<f',struct.pack('<I',b&0xFFFFFFFF))[0])
def seq(s,d,T):
    k=0; last=None
    while s<T:
        last=s; s=f32(s+d); k+=1
    return s,k,last
def jump(s,d,T):
    k=0; last=None; real=0
    while s<T:
        last=s; s=f32(s+d); k+=1; real+=1
        if not (s<T): break
        s2=f32(s+d); q=f32(s2-s)
        eb=bits(s)&0x7F800000
        top=frombits(eb+0x00800000)
        hu=frombits(eb-(24<<23))
        if s2<top and (f32(abs(f32(d-q)))!=hu or (bits(s)&1)==0):
            hi=min(T,top)
            est=f32(f32(hi-s)/q)*f32(0.99999)   # fdividef approx
            jf=f32(np.floor(est))
            sj=f32(np.float64(jf)*np.float64(q)+np.float64(s))  # fma exact in f64 then round
            n=0
            while f32(sj+q)<hi:
                sj=f32(sj+q); jf=f32(jf+1); n+=1
            assert n<=2,(n,s,d,T)
            s=sj; k+=int(jf)
    return s,k,last,real
random.seed(1)
tot_real=0; tot_k=0
hist={}
for it in range(300000):
    d=f32(10**random.uniform(-3,3))
    if random.random()<0.3:
        # force tie-prone d: few mantissa bits
        b=bits(d)&~((1<<random.randint(0,22))-1); d=frombits(b)
    s=f32(random.uniform(0,1)*float(d)) if random.random()<0.5 else f32(float(d)*random.uniform(0,3000))
    T=f32(float(s)+float(d)*random.uniform(0,5000)*random.random()**3)
    a=seq(s,d,T); b=jump(s,d,T)
    assert a[0]==b[0] and a[1]==b[1] and (a[2]==b[2]), (s,d,T,a,b)
    tot_real+=b[3]; tot_k+=b[1]; hist[b[3]]=hist.get(b[3],0)+1
print("ok; steps",tot_k,"real adds",tot_real)

print(sorted(hist.items())[:20], max(hist))
