// xn_kernels.cu -- the five volume-traversal kernels for sm_100a, plus the volume
// re-layout and synthetic-volume kernels.  Compiled with -fmad=false (see xn_device.cuh).
//
// Each kernel restates, from scratch, what one reference compute shader computes:
//   dda_kernel        resources/dda.comp        Amanatides-Woo grid march (multi-axis tie steps)
//   svo_naive_kernel  resources/svo_naive.comp  root-restart point location per leaf
//   svo_df_kernel     resources/svo_df.comp     exhaustive depth-first visit, explicit stack
//   esvo_kernel       resources/esvo.comp       Laine-Karras ESVO with emission accumulation
//   svo_rope_kernel   resources/svo_rope.comp   rope-tree leaf-to-leaf walk
// One thread per pixel; a warp owns an 8x4 pixel tile.  Traversal stacks live in shared
// memory ([level][thread], conflict-free) and are sized by the tree's real depth.
#include "xn_device.cuh"
#include "xn_kernels.h"
#include "xn_synth.h"

namespace xn {

template <bool STATS>
struct RayStats {
    uint32_t steps = 0;
    unsigned long long bytes = 0;
    __device__ __forceinline__ void step() {
        if (STATS) ++steps;
    }
    __device__ __forceinline__ void read(uint32_t n) {
        if (STATS) bytes += n;
    }
};

template <bool STATS>
__device__ __forceinline__ void store_result(const FrameParams& p, uint32_t ix, uint32_t iy, f3 color,
                                             const RayStats<STATS>& st) {
    p.target[(uint64_t)iy * p.target_stride + ix] = pack_pixel(color);
    if (STATS) {
        const uint64_t i = (uint64_t)iy * p.out_w + ix;
        if (p.steps_out) p.steps_out[i] = st.steps;
        if (p.bytes_out) p.bytes_out[i] = st.bytes;
    }
}

// ---------------------------------------------------------------------------------
// DDA (resources/dda.comp:13-73)
// ---------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(BLOCK_THREADS) dda_kernel(const __grid_constant__ FrameParams p) {
    uint32_t ix, iy;
    thread_pixel(p, ix, iy);
    if (ix >= p.out_w || iy >= p.out_h) return;
    RayStats<STATS> st;

    const f3 rd = make_ray(p, p.out_x + (int32_t)ix, p.out_y + (int32_t)iy);
    // textureSize(model) = grid dimensions; side = largest
    const float side = gmax((float)p.nx, gmax((float)p.ny, (float)p.nz));
    f3 ro = F3(p.pos[0] * side, p.pos[1] * side, p.pos[2] * side);
    const float ec = voxel_emission_coeff(p, rd) / side;

    const f3 rrd = F3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
    const f3 bias = F3(rrd.x * ro.x, rrd.y * ro.y, rrd.z * ro.z);
    const f3 bmin = F3(-bias.x, -bias.y, -bias.z);
    const f3 bmax = F3((float)p.model_dim[0] * rrd.x - bias.x, (float)p.model_dim[1] * rrd.y - bias.y,
                       (float)p.model_dim[2] * rrd.z - bias.z);
    float t_min = max_elem(F3(gmin(bmin.x, bmax.x), gmin(bmin.y, bmax.y), gmin(bmin.z, bmax.z)));
    const float t_max = min_elem(F3(gmax(bmin.x, bmax.x), gmax(bmin.y, bmax.y), gmax(bmin.z, bmax.z)));

    f3 total = F3(0.f, 0.f, 0.f);
    if (!(t_min > t_max)) {
        t_min = gmax(t_min, 0.0f);
        ro = F3(ro.x + rd.x * t_min, ro.y + rd.y * t_min, ro.z + rd.z * t_min);
        int px = (int)ro.x, py = (int)ro.y, pz = (int)ro.z; // ivec3(ro): truncation

        const f3 td = F3(fabsf(rrd.x), fabsf(rrd.y), fabsf(rrd.z));
        const f3 sg = F3(gsign(rd.x), gsign(rd.y), gsign(rd.z));
        const int sx = (int)sg.x, sy = (int)sg.y, sz = (int)sg.z;
        float sdx = (sg.x * ((floorf(ro.x) - ro.x) + 0.5f) + 0.5f) * td.x;
        float sdy = (sg.y * ((floorf(ro.y) - ro.y) + 0.5f) + 0.5f) * td.y;
        float sdz = (sg.z * ((floorf(ro.z) - ro.z) + 0.5f) + 0.5f) * td.z;

        const int64_t stride_y = (int64_t)p.nx, stride_z = (int64_t)p.nx * (int64_t)p.ny;
        const int64_t dix = sx, diy = sy * stride_y, diz = sz * stride_z;
        int64_t idx = (int64_t)px + (int64_t)py * stride_y + (int64_t)pz * stride_z;

        float t = 0.0f;
        const float t_end = t_max - t_min;
        while (t < t_end) {
            const bool mx = sdx <= fminf(sdy, sdz);
            const bool my = sdy <= fminf(sdz, sdx);
            const bool mz = sdz <= fminf(sdx, sdy);
            const float t0 = fminf(sdx, fminf(sdy, sdz));
            const float dt = t0 - t;

            // texelFetch; outside the grid -> 0 (border)
            if ((uint32_t)px < p.nx && (uint32_t)py < p.ny && (uint32_t)pz < p.nz) {
                const uint32_t v = __ldg(p.grid + idx);
                total.x += unorm8(v & 0xFFu) * dt;
                total.y += unorm8((v >> 8) & 0xFFu) * dt;
                total.z += unorm8((v >> 16) & 0xFFu) * dt;
            }
            t = t0;
            if (mx) { sdx += td.x; px += sx; idx += dix; }
            if (my) { sdy += td.y; py += sy; idx += diy; }
            if (mz) { sdz += td.z; pz += sz; idx += diz; }
            st.step();
            st.read(4);
        }
    }
    store_result(p, ix, iy, F3(total.x * ec, total.y * ec, total.z * ec), st);
}

// ---------------------------------------------------------------------------------
// shared octree helpers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ f3 meta_rgb(uint32_t m) {
    return F3(unorm8(m & 0xFFu), unorm8((m >> 8) & 0xFFu), unorm8((m >> 16) & 0xFFu));
}
__device__ __forceinline__ uint2 load_slot(const DNode* nodes, uint32_t node, uint32_t child) {
    return __ldg(&nodes[node].slot[child]);
}

// slab test of the unit cube [0,1]^3 (svo_naive.comp:30-45, svo_rope.comp:72-86)
__device__ __forceinline__ bool unit_cube_slab(f3 rrd, f3 bias, float& t_min, float& t_max) {
    const f3 bmin = F3(-bias.x, -bias.y, -bias.z);
    const f3 bmax = F3(rrd.x - bias.x, rrd.y - bias.y, rrd.z - bias.z);
    t_min = max_elem(F3(gmin(bmin.x, bmax.x), gmin(bmin.y, bmax.y), gmin(bmin.z, bmax.z)));
    t_max = min_elem(F3(gmax(bmin.x, bmax.x), gmax(bmin.y, bmax.y), gmax(bmin.z, bmax.z)));
    if (t_min > t_max) return false;
    t_min = gmax(t_min, 0.0f);
    return true;
}

// descend from (node, meta) at `offset`/`extent` to the leaf containing pos
// (loop body of find(), svo_naive.comp:14-26 == svo_rope.comp:14-26 == svo_rope.comp:33-47)
template <bool STATS>
__device__ __forceinline__ void descend(const DNode* nodes, f3 pos, uint32_t& node, uint32_t& meta, f3& offset,
                                        float& extent, RayStats<STATS>& st) {
    for (;;) {
        st.read(4); // is_leaf_depth
        if (meta_is_leaf(meta)) return;
        extent *= 0.5f;
        const bool mx = pos.x >= offset.x + extent;
        const bool my = pos.y >= offset.y + extent;
        const bool mz = pos.z >= offset.z + extent;
        const uint32_t child = (mx ? 4u : 0u) + (my ? 2u : 0u) + (mz ? 1u : 0u);
        offset.x += (mx ? 1.0f : 0.0f) * extent;
        offset.y += (my ? 1.0f : 0.0f) * extent;
        offset.z += (mz ? 1.0f : 0.0f) * extent;
        st.read(4); // children[child]
        const uint2 s = load_slot(nodes, node, child);
        node = s.x;
        meta = s.y;
    }
}

// chord of the ray through the node box [offset, offset+side] (svo_naive.comp:56-60)
__device__ __forceinline__ void node_slab(f3 offset, float side, f3 rrd, f3 bias, float& u_min, float& u_max,
                                          f3& far) {
    const f3 nmin = F3(offset.x * rrd.x - bias.x, offset.y * rrd.y - bias.y, offset.z * rrd.z - bias.z);
    const f3 nmax = F3((offset.x + side) * rrd.x - bias.x, (offset.y + side) * rrd.y - bias.y,
                       (offset.z + side) * rrd.z - bias.z);
    far = F3(gmax(nmin.x, nmax.x), gmax(nmin.y, nmax.y), gmax(nmin.z, nmax.z));
    u_min = max_elem(F3(gmin(nmin.x, nmax.x), gmin(nmin.y, nmax.y), gmin(nmin.z, nmax.z)));
    u_max = min_elem(far);
}

// ---------------------------------------------------------------------------------
// svo_naive (resources/svo_naive.comp:29-89)
// ---------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(BLOCK_THREADS) svo_naive_kernel(const __grid_constant__ FrameParams p) {
    uint32_t ix, iy;
    thread_pixel(p, ix, iy);
    if (ix >= p.out_w || iy >= p.out_h) return;
    RayStats<STATS> st;
    const float MIN_STEP_SIZE = 0.00001f;

    const f3 rd = make_ray(p, p.out_x + (int32_t)ix, p.out_y + (int32_t)iy);
    const f3 ro = F3(p.pos[0], p.pos[1], p.pos[2]);
    const f3 rrd = F3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
    const f3 bias = F3(rrd.x * ro.x, rrd.y * ro.y, rrd.z * ro.z);

    f3 total = F3(0.f, 0.f, 0.f);
    float t_min, t_max;
    if (unit_cube_slab(rrd, bias, t_min, t_max)) {
        float t = t_min + MIN_STEP_SIZE;
        while (t < t_max) {
            const f3 pt = F3(t * rd.x + ro.x, t * rd.y + ro.y, t * rd.z + ro.z);
            uint32_t node = 0, meta = p.root_meta;
            f3 offset = F3(0.f, 0.f, 0.f);
            float side = 1.0f;
            descend(p.nodes, pt, node, meta, offset, side, st);

            float u_min, u_max;
            f3 far;
            node_slab(offset, side, rrd, bias, u_min, u_max, far);
            u_min = gmax(u_min, 0.0f);
            const float step = gmax(u_max - u_min, MIN_STEP_SIZE);
            t += step;

            st.read(4); // color
            const f3 c = meta_rgb(meta);
            total.x += c.x * step;
            total.y += c.y * step;
            total.z += c.z * step;
            st.step();
        }
    }
    const float ec = voxel_emission_coeff(p, rd);
    store_result(p, ix, iy, F3(total.x * ec, total.y * ec, total.z * ec), st);
}

// ---------------------------------------------------------------------------------
// svo_df (resources/svo_df.comp:6-85)
// Stack entry = (node, child_idx | depth << 3): the depth of the pushed node is kept so
// the pop does not re-read is_leaf_depth (svo_df.comp:58) from memory.
// ---------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(BLOCK_THREADS) svo_df_kernel(const __grid_constant__ FrameParams p) {
    extern __shared__ uint2 df_stack[]; // [level][thread]
    uint32_t ix, iy;
    thread_pixel(p, ix, iy);
    if (ix >= p.out_w || iy >= p.out_h) return;
    RayStats<STATS> st;

    const f3 rd = make_ray(p, p.out_x + (int32_t)ix, p.out_y + (int32_t)iy);
    const f3 ro = F3(p.pos[0], p.pos[1], p.pos[2]);
    const f3 rrd = F3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
    const f3 bias = F3(rrd.x * ro.x, rrd.y * ro.y, rrd.z * ro.z);

    int sp = 0;
    uint32_t node = 0, child_idx = 0, depth = 0; // depth of `node`
    f3 pos = F3(0.f, 0.f, 0.f);
    float side = 0.5f;
    f3 total = F3(0.f, 0.f, 0.f);
    uint2* stack = df_stack + threadIdx.x;

    for (;;) {
        st.step();
        st.read(4); // children[child_idx]
        const uint2 s = load_slot(p.nodes, node, child_idx);
        const f3 bmin = F3(pos.x * rrd.x - bias.x, pos.y * rrd.y - bias.y, pos.z * rrd.z - bias.z);
        const f3 bmax = F3((pos.x + side) * rrd.x - bias.x, (pos.y + side) * rrd.y - bias.y,
                           (pos.z + side) * rrd.z - bias.z);
        const float t_min = max_elem(F3(gmin(bmin.x, bmax.x), gmin(bmin.y, bmax.y), gmin(bmin.z, bmax.z)));
        const float t_max = min_elem(F3(gmax(bmin.x, bmax.x), gmax(bmin.y, bmax.y), gmax(bmin.z, bmax.z)));

        if (t_min < t_max && t_max > 0.0f) {
            st.read(4); // nodes[child].is_leaf_depth
            if (meta_is_leaf(s.y)) {
                st.read(4); // color
                const f3 c = meta_rgb(s.y);
                const float len = t_max - gmax(t_min, 0.0f);
                total.x += c.x * len;
                total.y += c.y * len;
                total.z += c.z * len;
            } else {
                if (child_idx != 7u) {
                    stack[sp * BLOCK_THREADS] = make_uint2(node, child_idx | (depth << 3));
                    ++sp;
                }
                side *= 0.5f;
                node = s.x;
                depth = meta_depth(s.y);
                child_idx = 0;
                continue;
            }
        }

        if (child_idx == 7u) {
            --sp;
            if (sp < 0) break;
            const uint2 e = stack[sp * BLOCK_THREADS];
            node = e.x;
            child_idx = e.y & 7u;
            depth = e.y >> 3;
            st.read(4); // nodes[node].is_leaf_depth (svo_df.comp:58)
            side = __int_as_float((127 - (int)depth) << 23) * 0.5f; // exp2(-depth) * 0.5
        }

        const float s2 = side * 2.0f;
        pos.x -= gmod(pos.x, s2);
        pos.y -= gmod(pos.y, s2);
        pos.z -= gmod(pos.z, s2);
        ++child_idx;
        pos.x += (child_idx & 4u) ? side : 0.0f;
        pos.y += (child_idx & 2u) ? side : 0.0f;
        pos.z += (child_idx & 1u) ? side : 0.0f;
    }
    const float ec = voxel_emission_coeff(p, rd);
    store_result(p, ix, iy, F3(total.x * ec, total.y * ec, total.z * ec), st);
}

// ---------------------------------------------------------------------------------
// esvo (resources/esvo.comp:10-157)
// The reference indexes its stacks by `scale` (22 downwards); here level = 22 - scale,
// so only max_depth + 1 levels of shared memory are needed.
// ---------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(BLOCK_THREADS) esvo_kernel(const __grid_constant__ FrameParams p) {
    extern __shared__ uint2 esvo_stack[]; // [level][thread] = (parent, bits(t_max))
    uint32_t ix, iy;
    thread_pixel(p, ix, iy);
    if (ix >= p.out_w || iy >= p.out_h) return;
    RayStats<STATS> st;
    const uint32_t cast_stack_depth = 23u;

    const f3 rd = make_ray(p, p.out_x + (int32_t)ix, p.out_y + (int32_t)iy);
    f3 ro = F3(p.pos[0] + 1.0f, p.pos[1] + 1.0f, p.pos[2] + 1.0f);
    {
        // aabb_intersect(vec3(1), vec3(2), ro, rd), esvo.comp:10-21
        const f3 r = F3(1.0f / (rd.x + 0.00000001f), 1.0f / (rd.y + 0.00000001f), 1.0f / (rd.z + 0.00000001f));
        const f3 tbot = F3((1.0f - ro.x) * r.x, (1.0f - ro.y) * r.y, (1.0f - ro.z) * r.z);
        const f3 ttop = F3((2.0f - ro.x) * r.x, (2.0f - ro.y) * r.y, (2.0f - ro.z) * r.z);
        const f3 tmn = F3(gmin(ttop.x, tbot.x), gmin(ttop.y, tbot.y), gmin(ttop.z, tbot.z));
        const float t0 = gmax(gmax(tmn.x, tmn.y), gmax(tmn.x, tmn.z));
        const float adv = gmax(t0, 0.0f);
        ro = F3(ro.x + adv * rd.x, ro.y + adv * rd.y, ro.z + adv * rd.z);
    }

    const f3 tc = F3(1.0f / -fabsf(rd.x), 1.0f / -fabsf(rd.y), 1.0f / -fabsf(rd.z));
    f3 tb = F3(tc.x * ro.x, tc.y * ro.y, tc.z * ro.z);
    uint32_t octant_mask = 0;
    if (rd.x > 0.0f) { tb.x = 3.0f * tc.x - tb.x; octant_mask ^= 4u; }
    if (rd.y > 0.0f) { tb.y = 3.0f * tc.y - tb.y; octant_mask ^= 2u; }
    if (rd.z > 0.0f) { tb.z = 3.0f * tc.z - tb.z; octant_mask ^= 1u; }

    float t_min = max_elem(F3(2.0f * tc.x - tb.x, 2.0f * tc.y - tb.y, 2.0f * tc.z - tb.z));
    float t_max = min_elem(F3(tc.x - tb.x, tc.y - tb.y, tc.z - tb.z));
    float h = t_max;
    t_min = gmax(t_min, 0.0f);
    t_max = gmin(t_max, sqrtf(3.0f));

    uint32_t parent = 0, idx = 0;
    f3 pos = F3(1.f, 1.f, 1.f);
    uint32_t scale = cast_stack_depth - 1u;
    float scale_exp2 = 0.5f;
    if (1.5f * tc.x - tb.x > t_min) { pos.x = 1.5f; idx ^= 4u; }
    if (1.5f * tc.y - tb.y > t_min) { pos.y = 1.5f; idx ^= 2u; }
    if (1.5f * tc.z - tb.z > t_min) { pos.z = 1.5f; idx ^= 1u; }

    f3 total = F3(0.f, 0.f, 0.f);
    uint2* stack = esvo_stack + threadIdx.x;
    const uint32_t levels = p.max_depth + 1u;

    while (scale < cast_stack_depth) {
        st.step();
        const f3 t_corner = F3(pos.x * tc.x - tb.x, pos.y * tc.y - tb.y, pos.z * tc.z - tb.z);
        const float tc_max = min_elem(t_corner);

        if (t_min <= t_max) {
            const float tv_max = gmin(t_max, tc_max);
            if (t_min <= tv_max) {
                st.read(8); // children[idx ^ octant_mask] + nodes[child].is_leaf_depth
                const uint2 s = load_slot(p.nodes, parent, idx ^ octant_mask);
                if (meta_is_leaf(s.y)) {
                    st.read(4); // color
                    const f3 c = meta_rgb(s.y);
                    const float len = tv_max - t_min;
                    total.x += c.x * len;
                    total.y += c.y * len;
                    total.z += c.z * len;
                } else {
                    // PUSH
                    if (tc_max < h) {
                        const uint32_t level = (cast_stack_depth - 1u) - scale;
                        if (level < levels) stack[level * BLOCK_THREADS] = make_uint2(parent, __float_as_uint(t_max));
                    }
                    h = tc_max;
                    parent = s.x;
                    --scale;
                    scale_exp2 *= 0.5f;
                    const f3 t_center = F3(scale_exp2 * tc.x + t_corner.x, scale_exp2 * tc.y + t_corner.y,
                                           scale_exp2 * tc.z + t_corner.z);
                    idx = 0;
                    if (t_center.x > t_min) { idx ^= 4u; pos.x += scale_exp2; }
                    if (t_center.y > t_min) { idx ^= 2u; pos.y += scale_exp2; }
                    if (t_center.z > t_min) { idx ^= 1u; pos.z += scale_exp2; }
                    t_max = tv_max;
                    continue;
                }
            }
        }

        // ADVANCE
        const bool ax = t_corner.x <= tc_max, ay = t_corner.y <= tc_max, az = t_corner.z <= tc_max;
        const uint32_t step_mask = (ax ? 4u : 0u) ^ (ay ? 2u : 0u) ^ (az ? 1u : 0u);
        if (ax) pos.x -= scale_exp2;
        if (ay) pos.y -= scale_exp2;
        if (az) pos.z -= scale_exp2;
        t_min = tc_max;
        idx ^= step_mask;

        if ((idx & step_mask) != 0u) {
            // POP
            uint32_t dbits = 0;
            if (ax) dbits |= __float_as_uint(pos.x) ^ __float_as_uint(pos.x + scale_exp2);
            if (ay) dbits |= __float_as_uint(pos.y) ^ __float_as_uint(pos.y + scale_exp2);
            if (az) dbits |= __float_as_uint(pos.z) ^ __float_as_uint(pos.z + scale_exp2);
            scale = (__float_as_uint((float)dbits) >> 23) - 127u;
            if (scale >= cast_stack_depth) break; // left the cube (also guards the reference's
                                                  // underflowed stack read, esvo.comp:119-123)
            scale_exp2 = __uint_as_float((scale - cast_stack_depth + 127u) << 23);
            const uint32_t level = (cast_stack_depth - 1u) - scale;
            const uint2 e = level < levels ? stack[level * BLOCK_THREADS] : make_uint2(0u, 0u);
            parent = e.x;
            t_max = __uint_as_float(e.y);
            const uint32_t shx = __float_as_uint(pos.x) >> scale, shy = __float_as_uint(pos.y) >> scale,
                           shz = __float_as_uint(pos.z) >> scale;
            pos.x = __uint_as_float(shx << scale);
            pos.y = __uint_as_float(shy << scale);
            pos.z = __uint_as_float(shz << scale);
            idx = (shx & 1u) * 4u + (shy & 1u) * 2u + (shz & 1u);
            h = 0.0f;
        }
    }
    const float ec = voxel_emission_coeff(p, rd);
    store_result(p, ix, iy, F3(total.x * ec, total.y * ec, total.z * ec), st);
}

// ---------------------------------------------------------------------------------
// svo_rope (resources/svo_rope.comp:50-153)
// ---------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(BLOCK_THREADS) svo_rope_kernel(const __grid_constant__ FrameParams p) {
    uint32_t ix, iy;
    thread_pixel(p, ix, iy);
    if (ix >= p.out_w || iy >= p.out_h) return;
    RayStats<STATS> st;

    const f3 rd = make_ray(p, p.out_x + (int32_t)ix, p.out_y + (int32_t)iy);
    const f3 ro = F3(p.pos[0], p.pos[1], p.pos[2]);

    f3 sgn = F3(gsign(rd.x), gsign(rd.y), gsign(rd.z));
    const uint32_t nbx = 1u - (uint32_t)gmax(sgn.x, 0.0f);
    const uint32_t nby = 3u - (uint32_t)gmax(sgn.y, 0.0f);
    const uint32_t nbz = 5u - (uint32_t)gmax(sgn.z, 0.0f);
    sgn = F3(sgn.x + 0.1f, sgn.y + 0.1f, sgn.z + 0.1f);

    const f3 rrd = F3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
    const f3 bias = F3(rrd.x * ro.x, rrd.y * ro.y, rrd.z * ro.z);

    f3 total = F3(0.f, 0.f, 0.f);
    float t_min, t_max;
    if (unit_cube_slab(rrd, bias, t_min, t_max)) {
        f3 pos = F3(ro.x + t_min * rd.x, ro.y + t_min * rd.y, ro.z + t_min * rd.z);
        uint32_t node = 0, meta = p.root_meta;
        f3 offset = F3(0.f, 0.f, 0.f);
        float side = 1.0f;
        descend(p.nodes, pos, node, meta, offset, side, st);

        bool first = true;
        for (;;) {
            float u_min, u_max;
            f3 far;
            node_slab(offset, side, rrd, bias, u_min, u_max, far);
            const float step = u_max - gmax(u_min, 0.0f);
            st.read(4); // color
            const f3 c = meta_rgb(meta);
            if (first) {
                total = F3(c.x * step, c.y * step, c.z * step);
                first = false;
            } else {
                total.x += c.x * step;
                total.y += c.y * step;
                total.z += c.z * step;
            }
            st.step();

            // neighbor_index, svo_rope.comp:50-63 (ties go to the later axis)
            uint32_t n;
            if (far.x < gmin(far.y, far.z)) {
                n = nbx;
                offset.x += sgn.x * side;
            } else if (far.y < far.z) {
                n = nby;
                offset.y += sgn.y * side;
            } else {
                n = nbz;
                offset.z += sgn.z * side;
            }
            st.read(4); // rope
            const uint2 s = load_slot(p.nodes, node, n);
            node = s.x;
            meta = s.y;
            if (node == 0u) break;

            // find_relative, svo_rope.comp:29-48
            pos = F3(ro.x + u_max * rd.x, ro.y + u_max * rd.y, ro.z + u_max * rd.z);
            side = __int_as_float((127 - (int)meta_depth(meta)) << 23); // exp2(-depth)
            offset.x = offset.x - gmod(offset.x, side);
            offset.y = offset.y - gmod(offset.y, side);
            offset.z = offset.z - gmod(offset.z, side);
            descend(p.nodes, pos, node, meta, offset, side, st);
        }
    }
    const float ec = voxel_emission_coeff(p, rd);
    store_result(p, ix, iy, F3(total.x * ec, total.y * ec, total.z * ec), st);
}

// ---------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------
template <bool STATS>
static cudaError_t launch_t(int traversal, const FrameParams& p, cudaStream_t stream) {
    const uint32_t stripes = (p.out_h + BLOCK_H - 1) / BLOCK_H;
    if (p.il_count == 0 || p.il_index >= p.il_count) return cudaErrorInvalidValue;
    const uint32_t owned = stripes > p.il_index ? (stripes - p.il_index + p.il_count - 1) / p.il_count : 0;
    if (owned == 0) return cudaSuccess;
    const dim3 grid((p.out_w + BLOCK_W - 1) / BLOCK_W, owned, 1);
    const dim3 block(BLOCK_THREADS, 1, 1);
    const size_t stack_bytes = (size_t)(p.max_depth + 1u) * BLOCK_THREADS * sizeof(uint2);
    switch (traversal) {
        case 0: dda_kernel<STATS><<<grid, block, 0, stream>>>(p); break;
        case 1: svo_naive_kernel<STATS><<<grid, block, 0, stream>>>(p); break;
        case 2: esvo_kernel<STATS><<<grid, block, stack_bytes, stream>>>(p); break;
        case 3: svo_df_kernel<STATS><<<grid, block, stack_bytes, stream>>>(p); break;
        case 4: svo_rope_kernel<STATS><<<grid, block, 0, stream>>>(p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_traversal(int traversal, const FrameParams& p, bool stats, cudaStream_t stream) {
    if (p.out_w == 0 || p.out_h == 0) return cudaSuccess;
    return stats ? launch_t<true>(traversal, p, stream) : launch_t<false>(traversal, p, stream);
}

cudaError_t configure_kernels() {
    // stacks of up to 24 levels x 256 threads x 8 B = 48 KB can exceed the default limit
    const int max_stack = 24 * BLOCK_THREADS * (int)sizeof(uint2);
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(esvo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_stack))) return e;
    if ((e = cudaFuncSetAttribute(esvo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_stack))) return e;
    if ((e = cudaFuncSetAttribute(svo_df_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_stack))) return e;
    if ((e = cudaFuncSetAttribute(svo_df_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_stack))) return e;
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------
// volume re-layout: 40-byte file nodes -> 64-byte device nodes (one thread per node)
// ---------------------------------------------------------------------------------
__global__ void relayout_nodes_kernel(const uint32_t* __restrict__ raw, uint64_t count, DNode* __restrict__ out,
                                      uint32_t* __restrict__ max_depth) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t my_depth = 0;
    if (i < count) {
        const uint32_t* n = raw + i * 10u;
        my_depth = n[9] & 0x7FFFFFFFu;
        DNode d;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t child = n[c];
            if (child >= count) child = 0; // malformed file: never index out of bounds
            const uint32_t* cn = raw + (uint64_t)child * 10u;
            d.slot[c] = make_uint2(child, make_meta(cn[8], cn[9]));
        }
        uint4* o = reinterpret_cast<uint4*>(out + i);
        o[0] = make_uint4(d.slot[0].x, d.slot[0].y, d.slot[1].x, d.slot[1].y);
        o[1] = make_uint4(d.slot[2].x, d.slot[2].y, d.slot[3].x, d.slot[3].y);
        o[2] = make_uint4(d.slot[4].x, d.slot[4].y, d.slot[5].x, d.slot[5].y);
        o[3] = make_uint4(d.slot[6].x, d.slot[6].y, d.slot[7].x, d.slot[7].y);
    }
    // block-wide max of depth -> one atomic per warp
    for (int o = 16; o > 0; o >>= 1) my_depth = max(my_depth, __shfl_xor_sync(0xFFFFFFFFu, my_depth, o));
    if ((threadIdx.x & 31) == 0 && my_depth) atomicMax(max_depth, my_depth);
}

cudaError_t launch_relayout(const void* raw40, uint64_t count, DNode* out, uint32_t* d_max_depth,
                            cudaStream_t stream) {
    const int threads = 256;
    const uint64_t blocks = (count + threads - 1) / threads;
    relayout_nodes_kernel<<<(unsigned)blocks, threads, 0, stream>>>((const uint32_t*)raw40, count, out, d_max_depth);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// synthetic volumes (bit-identical to the host generator in xn_synth.h)
// ---------------------------------------------------------------------------------
__global__ void synth_kernel(uint32_t* __restrict__ grid, SynthSpec spec) {
    const uint64_t n = (uint64_t)spec.nx * spec.ny * spec.nz;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = (uint32_t)(i % spec.nx);
        const uint32_t y = (uint32_t)((i / spec.nx) % spec.ny);
        const uint32_t z = (uint32_t)(i / ((uint64_t)spec.nx * spec.ny));
        grid[i] = synth_voxel(spec, x, y, z);
    }
}

cudaError_t launch_synth(uint32_t* grid, const SynthSpec& spec, cudaStream_t stream) {
    synth_kernel<<<148 * 16, 256, 0, stream>>>(grid, spec);
    return cudaGetLastError();
}

// sum reduction of the stats arrays (totals for the roofline accounting)
__global__ void stats_totals_kernel(const uint32_t* __restrict__ steps, const unsigned long long* __restrict__ bytes,
                                    uint64_t n, unsigned long long* __restrict__ totals) {
    unsigned long long s = 0, b = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        s += steps[i];
        b += bytes[i];
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&totals[0], s);
        atomicAdd(&totals[1], b);
    }
}

cudaError_t launch_stats_totals(const uint32_t* steps, const unsigned long long* bytes, uint64_t n,
                                unsigned long long* totals, cudaStream_t stream) {
    stats_totals_kernel<<<148 * 4, 256, 0, stream>>>(steps, bytes, n, totals);
    return cudaGetLastError();
}

} // namespace xn
