/*
 * xn_oracle.c -- CPU restatement of the reference's five traversal shaders and
 * of `xenodon convert`.  TEST INFRASTRUCTURE ONLY (see xn_oracle.h).
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference).  Arithmetic is binary32 in source order; build with
 * -ffp-contract=off (oracle/Makefile) so no FMA is formed.
 */
#include "xn_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* GLSL built-ins, as used by the shaders       */
/* ------------------------------------------------------------------ */

typedef struct { float x, y, z; } v3;

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3s(float s) { return V3(s, s, s); }
static inline v3 vadd(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vdiv(v3 a, v3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 vscale(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline float fmin2(float a, float b) { return b < a ? b : a; } /* GLSL min */
static inline float fmax2(float a, float b) { return a < b ? b : a; } /* GLSL max */
static inline v3 vmin(v3 a, v3 b) { return V3(fmin2(a.x, b.x), fmin2(a.y, b.y), fmin2(a.z, b.z)); }
static inline v3 vmax(v3 a, v3 b) { return V3(fmax2(a.x, b.x), fmax2(a.y, b.y), fmax2(a.z, b.z)); }
static inline float min_elem(v3 v) { return fmin2(v.x, fmin2(v.y, v.z)); } /* common.glsl:58-60 */
static inline float max_elem(v3 v) { return fmax2(v.x, fmax2(v.y, v.z)); } /* common.glsl:62-64 */
static inline float dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 cross3(v3 a, v3 b) {
    return V3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline v3 normalize3(v3 a) {
    float len = sqrtf(dot3(a, a));
    return V3(a.x / len, a.y / len, a.z / len);
}
static inline float sign1(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
static inline float mod1(float x, float y) { return x - y * floorf(x / y); }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float exp2_neg_int(uint32_t d) { return ldexpf(1.0f, -(int)d); }

/* unpackUnorm4x8(u).rgb */
static inline v3 unpack_rgb(uint32_t c) {
    return V3((float)(c & 0xFF) / 255.0f, (float)((c >> 8) & 0xFF) / 255.0f,
              (float)((c >> 16) & 0xFF) / 255.0f);
}

#define LEAF_MASK 0x80000000u   /* resources/octree.glsl:16 */
#define DEPTH_MASK 0x7FFFFFFFu  /* resources/octree.glsl:17 */
#define FLOAT_MANTISSA_BITS 23u /* resources/common.glsl:37 */

typedef struct {
    const xo_volume* vol;
    xo_params p;
    v3 ratio;
    v3 fwd, up, translation; /* translation already divided by ratio */
    xo_rect out, disp;
} frame_ctx;

typedef struct {
    uint32_t steps;
    uint64_t bytes;
} ray_stats;

/* resources/common.glsl:40-43 */
static v3 adjust_ray(v3 rd) {
    const float epsilon = 1.1920928955078125e-07f; /* exp2(-23) */
    if (fabsf(rd.x) < epsilon) rd.x = epsilon;
    if (fabsf(rd.y) < epsilon) rd.y = epsilon;
    if (fabsf(rd.z) < epsilon) rd.z = epsilon;
    return rd;
}

/* resources/common.glsl:45-56 */
static v3 ray(const frame_ctx* c, float uvx, float uvy) {
    uvx -= 0.5f;
    uvy -= 0.5f;
    uvy *= (float)c->disp.h / (float)c->disp.w;

    v3 dir = c->fwd;
    v3 up = c->up;
    v3 right = normalize3(cross3(up, dir));
    up = normalize3(cross3(right, dir));

    v3 rd = normalize3(vadd(vadd(vscale(right, uvx), vscale(up, uvy)), dir));
    return adjust_ray(normalize3(vdiv(rd, c->ratio)));
}

/* resources/common.glsl:68-72 */
static float voxel_emission_coeff(const frame_ctx* c, v3 rd) {
    v3 rd2 = vmul(rd, rd);
    v3 dim2 = vmul(c->ratio, c->ratio);
    return c->p.emission_coeff * sqrtf(dot3(rd2, dim2) / dot3(rd2, v3s(1.0f)));
}

/* imageStore to an rgba8 image: clamp, scale, round to nearest even */
static uint32_t pack_unorm8(float v) {
    if (!(v > 0.0f)) return 0; /* also NaN -> 0 */
    if (v > 1.0f) v = 1.0f;
    return (uint32_t)rintf(v * 255.0f);
}

/* ------------------------------------------------------------------ */
/* dda.comp                                                            */
/* ------------------------------------------------------------------ */

/* texelFetch(model, p, 0).rgb; out of range -> 0 (Appendix A) */
static v3 dda_texel(const xo_volume* vol, int x, int y, int z) {
    if (x < 0 || y < 0 || z < 0 || (uint64_t)x >= vol->nx || (uint64_t)y >= vol->ny ||
        (uint64_t)z >= vol->nz)
        return v3s(0.0f);
    const uint8_t* px =
        vol->grid + 4 * ((uint64_t)x + (uint64_t)y * vol->nx + (uint64_t)z * vol->nx * vol->ny);
    return V3((float)px[0] / 255.0f, (float)px[1] / 255.0f, (float)px[2] / 255.0f);
}

/* resources/dda.comp:13-53 */
static v3 dda_trace(const frame_ctx* c, v3 ro, v3 rd, ray_stats* st) {
    v3 rrd = vdiv(v3s(1.0f), rd);
    v3 bias = vmul(rrd, ro);

    v3 dim = V3((float)c->p.model_dim[0], (float)c->p.model_dim[1], (float)c->p.model_dim[2]);
    v3 box_min = V3(-bias.x, -bias.y, -bias.z);
    v3 box_max = vsub(vmul(dim, rrd), bias);

    float t_min = max_elem(vmin(box_min, box_max));
    float t_max = min_elem(vmax(box_min, box_max));

    if (t_min > t_max) return v3s(0.0f);

    t_min = fmax2(t_min, 0.0f);

    ro = vadd(ro, vscale(rd, t_min));
    int px = (int)ro.x, py = (int)ro.y, pz = (int)ro.z; /* ivec3(ro): truncation */

    v3 t_delta = V3(fabsf(rrd.x), fabsf(rrd.y), fabsf(rrd.z));
    v3 sgn = V3(sign1(rd.x), sign1(rd.y), sign1(rd.z));
    int sx = (int)sgn.x, sy = (int)sgn.y, sz = (int)sgn.z;
    v3 fl = V3(floorf(ro.x), floorf(ro.y), floorf(ro.z));
    v3 side_dist =
        vmul(vadd(vmul(sgn, vadd(vsub(fl, ro), v3s(0.5f))), v3s(0.5f)), t_delta);

    float t = 0.0f;
    v3 total = v3s(0.0f);
    const float t_end = t_max - t_min;
    while (t < t_end) {
        int mx = side_dist.x <= fmin2(side_dist.y, side_dist.z);
        int my = side_dist.y <= fmin2(side_dist.z, side_dist.x);
        int mz = side_dist.z <= fmin2(side_dist.x, side_dist.y);

        float t0 = min_elem(side_dist);
        total = vadd(total, vscale(dda_texel(c->vol, px, py, pz), t0 - t));
        t = t0;

        side_dist.x += mx ? t_delta.x : 0.0f;
        side_dist.y += my ? t_delta.y : 0.0f;
        side_dist.z += mz ? t_delta.z : 0.0f;
        px += mx ? sx : 0;
        py += my ? sy : 0;
        pz += mz ? sz : 0;

        st->steps += 1;
        st->bytes += 4;
    }
    return total;
}

/* ------------------------------------------------------------------ */
/* svo_naive.comp                                                      */
/* ------------------------------------------------------------------ */

#define MIN_STEP_SIZE 0.00001f /* svo_naive.comp:6, svo_rope.comp:6 */

/* svo_naive.comp:8-27 and svo_rope.comp:8-27 (identical) */
static uint32_t svo_find(const xo_node* nodes, v3 pos, v3* base, float* side, ray_stats* st) {
    float extent = 1.0f;
    uint32_t index = 0;
    v3 offset = v3s(0.0f);
    for (;;) {
        st->bytes += 4;
        if (nodes[index].is_leaf_depth >= LEAF_MASK) {
            *base = offset;
            *side = extent;
            return index;
        }
        extent *= 0.5f;
        int mx = pos.x >= offset.x + extent;
        int my = pos.y >= offset.y + extent;
        int mz = pos.z >= offset.z + extent;
        int child = mx * 4 + my * 2 + mz;
        offset.x += (float)mx * extent;
        offset.y += (float)my * extent;
        offset.z += (float)mz * extent;
        st->bytes += 4;
        index = nodes[index].children[child];
    }
}

/* svo_naive.comp:29-71 */
static v3 svo_naive_trace(const frame_ctx* c, v3 ro, v3 rd, ray_stats* st) {
    const xo_node* nodes = c->vol->nodes;
    v3 rrd = vdiv(v3s(1.0f), rd);
    v3 bias = vmul(rrd, ro);
    v3 box_min = V3(-bias.x, -bias.y, -bias.z);
    v3 box_max = vsub(rrd, bias);
    float t_min = max_elem(vmin(box_min, box_max));
    float t_max = min_elem(vmax(box_min, box_max));
    if (t_min > t_max) return v3s(0.0f);
    t_min = fmax2(t_min, 0.0f);

    v3 total = v3s(0.0f);
    float t = t_min + MIN_STEP_SIZE;
    while (t < t_max) {
        v3 p = vadd(vscale(rd, t), ro);
        v3 offset;
        float side;
        uint32_t node = svo_find(nodes, p, &offset, &side, st);

        v3 node_min = vsub(vmul(offset, rrd), bias);
        v3 node_max = vsub(vmul(vadd(offset, v3s(side)), rrd), bias);
        float u_min = max_elem(vmin(node_min, node_max));
        float u_max = min_elem(vmax(node_min, node_max));
        u_min = fmax2(u_min, 0.0f);
        float step = fmax2(u_max - u_min, MIN_STEP_SIZE);
        t += step;

        st->bytes += 4;
        v3 color = unpack_rgb(nodes[node].color);
        total = vadd(total, vscale(color, step));
        st->steps += 1;
    }
    return total;
}

/* ------------------------------------------------------------------ */
/* svo_df.comp                                                         */
/* ------------------------------------------------------------------ */

/* svo_df.comp:6-67 */
static v3 svo_df_trace(const frame_ctx* c, v3 ro, v3 rd, ray_stats* st) {
    const xo_node* nodes = c->vol->nodes;
    v3 rrd = vdiv(v3s(1.0f), rd);
    v3 bias = vmul(rrd, ro);

    int sp = 0;
    uint32_t node_stack[FLOAT_MANTISSA_BITS];
    uint32_t child_index_stack[FLOAT_MANTISSA_BITS];

    uint32_t node = 0;
    uint32_t child_idx = 0;
    v3 pos = v3s(0.0f);
    float side = 0.5f;
    v3 total = v3s(0.0f);

    for (;;) {
        st->steps += 1;
        st->bytes += 4;
        uint32_t child = nodes[node].children[child_idx];
        v3 box_min = vsub(vmul(pos, rrd), bias);
        v3 box_max = vsub(vmul(vadd(pos, v3s(side)), rrd), bias);
        float t_min = max_elem(vmin(box_min, box_max));
        float t_max = min_elem(vmax(box_min, box_max));

        if (t_min < t_max && t_max > 0.0f) {
            st->bytes += 4;
            if (nodes[child].is_leaf_depth >= LEAF_MASK) {
                st->bytes += 4;
                v3 color = unpack_rgb(nodes[child].color);
                total = vadd(total, vscale(color, t_max - fmax2(t_min, 0.0f)));
            } else {
                if (child_idx != 7) {
                    if (sp < (int)FLOAT_MANTISSA_BITS) { /* guard: GLSL would be UB */
                        node_stack[sp] = node;
                        child_index_stack[sp] = child_idx;
                    }
                    ++sp;
                }
                side *= 0.5f;
                node = child;
                child_idx = 0;
                continue;
            }
        }

        if (child_idx == 7) {
            --sp;
            if (sp < 0) break;
            node = node_stack[sp];
            child_idx = child_index_stack[sp];
            st->bytes += 4;
            side = exp2_neg_int(nodes[node].is_leaf_depth & DEPTH_MASK) * 0.5f;
        }

        float s2 = side * 2.0f;
        pos.x -= mod1(pos.x, s2);
        pos.y -= mod1(pos.y, s2);
        pos.z -= mod1(pos.z, s2);
        ++child_idx;
        pos.x += (child_idx & 4u) ? side : 0.0f;
        pos.y += (child_idx & 2u) ? side : 0.0f;
        pos.z += (child_idx & 1u) ? side : 0.0f;
    }
    return total;
}

/* ------------------------------------------------------------------ */
/* esvo.comp                                                           */
/* ------------------------------------------------------------------ */

/* esvo.comp:10-21 */
static void aabb_intersect(v3 bmin, v3 bmax, v3 ro, v3 rd, float* t0o, float* t1o) {
    v3 rrd = vdiv(v3s(1.0f), vadd(rd, v3s(0.00000001f)));
    v3 tbot = vmul(vsub(bmin, ro), rrd);
    v3 ttop = vmul(vsub(bmax, ro), rrd);
    v3 tmin = vmin(ttop, tbot);
    v3 tmax = vmax(ttop, tbot);
    float tx = fmax2(tmin.x, tmin.y), ty = fmax2(tmin.x, tmin.z);
    *t0o = fmax2(tx, ty);
    tx = fmin2(tmax.x, tmax.y);
    ty = fmin2(tmax.x, tmax.z);
    *t1o = fmin2(tx, ty);
}

/* esvo.comp:23-28 */
static inline uint32_t cxor(uint32_t x, int ax, int ay, int az) {
    x ^= ax ? 4u : 0u;
    x ^= ay ? 2u : 0u;
    x ^= az ? 1u : 0u;
    return x;
}

/* esvo.comp:30-136 */
static v3 esvo_trace(const frame_ctx* c, v3 ro, v3 rd, ray_stats* st) {
    const xo_node* nodes = c->vol->nodes;
    const uint32_t cast_stack_depth = FLOAT_MANTISSA_BITS;
    uint32_t node_stack[FLOAT_MANTISSA_BITS + 1];
    float t_max_stack[FLOAT_MANTISSA_BITS + 1];
    memset(node_stack, 0, sizeof node_stack);
    memset(t_max_stack, 0, sizeof t_max_stack);

    v3 t_coeff = V3(1.0f / -fabsf(rd.x), 1.0f / -fabsf(rd.y), 1.0f / -fabsf(rd.z));
    v3 t_bias = vmul(t_coeff, ro);

    int gx = rd.x > 0.0f, gy = rd.y > 0.0f, gz = rd.z > 0.0f;
    if (gx) t_bias.x = 3.0f * t_coeff.x - t_bias.x;
    if (gy) t_bias.y = 3.0f * t_coeff.y - t_bias.y;
    if (gz) t_bias.z = 3.0f * t_coeff.z - t_bias.z;
    uint32_t octant_mask = cxor(0, gx, gy, gz);

    float t_min = max_elem(vsub(vscale(t_coeff, 2.0f), t_bias));
    float t_max = min_elem(vsub(t_coeff, t_bias));
    float h = t_max;

    t_min = fmax2(t_min, 0.0f);
    t_max = fmin2(t_max, sqrtf(3.0f));

    uint32_t parent = 0;
    uint32_t idx = 0;
    v3 pos = v3s(1.0f);
    uint32_t scale = cast_stack_depth - 1;
    float scale_exp2 = 0.5f;

    {
        int ax = 1.5f * t_coeff.x - t_bias.x > t_min;
        int ay = 1.5f * t_coeff.y - t_bias.y > t_min;
        int az = 1.5f * t_coeff.z - t_bias.z > t_min;
        if (ax) pos.x = 1.5f;
        if (ay) pos.y = 1.5f;
        if (az) pos.z = 1.5f;
        idx = cxor(idx, ax, ay, az);
    }

    v3 total = v3s(0.0f);

    while (scale < cast_stack_depth) {
        st->steps += 1;
        v3 t_corner = vsub(vmul(pos, t_coeff), t_bias);
        float tc_max = min_elem(t_corner);

        if (t_min <= t_max) {
            float tv_max = fmin2(t_max, tc_max);
            if (t_min <= tv_max) {
                st->bytes += 8;
                uint32_t child = nodes[parent].children[idx ^ octant_mask];
                if (nodes[child].is_leaf_depth >= LEAF_MASK) {
                    st->bytes += 4;
                    v3 color = unpack_rgb(nodes[child].color);
                    total = vadd(total, vscale(color, tv_max - t_min));
                } else {
                    /* PUSH */
                    if (tc_max < h) {
                        node_stack[scale] = parent;
                        t_max_stack[scale] = t_max;
                    }
                    h = tc_max;
                    parent = child;
                    --scale;
                    scale_exp2 *= 0.5f;

                    v3 t_center = vadd(vscale(t_coeff, scale_exp2), t_corner);
                    int ax = t_center.x > t_min, ay = t_center.y > t_min, az = t_center.z > t_min;
                    idx = cxor(0, ax, ay, az);
                    pos.x += ax ? scale_exp2 : 0.0f;
                    pos.y += ay ? scale_exp2 : 0.0f;
                    pos.z += az ? scale_exp2 : 0.0f;
                    t_max = tv_max;
                    continue;
                }
            }
        }

        /* ADVANCE */
        int ax = t_corner.x <= tc_max, ay = t_corner.y <= tc_max, az = t_corner.z <= tc_max;
        uint32_t step_mask = cxor(0, ax, ay, az);
        pos.x -= ax ? scale_exp2 : 0.0f;
        pos.y -= ay ? scale_exp2 : 0.0f;
        pos.z -= az ? scale_exp2 : 0.0f;

        t_min = tc_max;
        idx ^= step_mask;

        if ((idx & step_mask) != 0) {
            /* POP */
            uint32_t xx = f2u(pos.x) ^ f2u(pos.x + scale_exp2);
            uint32_t xy = f2u(pos.y) ^ f2u(pos.y + scale_exp2);
            uint32_t xz = f2u(pos.z) ^ f2u(pos.z + scale_exp2);
            uint32_t dbits = (ax ? xx : 0u) | (ay ? xy : 0u) | (az ? xz : 0u);

            scale = (f2u((float)dbits) >> 23) - 127u;
            if (scale >= cast_stack_depth) {
                /* left the cube (or dbits == 0: the underflow hazard of
                 * esvo.comp:119-123); the while condition ends the loop and
                 * nothing computed below is observable. */
                break;
            }
            scale_exp2 = u2f((scale - cast_stack_depth + 127u) << 23);

            parent = node_stack[scale];
            t_max = t_max_stack[scale];

            uint32_t shx = f2u(pos.x) >> scale, shy = f2u(pos.y) >> scale, shz = f2u(pos.z) >> scale;
            pos.x = u2f(shx << scale);
            pos.y = u2f(shy << scale);
            pos.z = u2f(shz << scale);
            idx = (shx & 1u) * 4u + (shy & 1u) * 2u + (shz & 1u);
            h = 0.0f;
        }
    }
    return total;
}

/* ------------------------------------------------------------------ */
/* svo_rope.comp                                                       */
/* ------------------------------------------------------------------ */

/* svo_rope.comp:29-48 */
static uint32_t svo_find_relative(const xo_node* nodes, uint32_t parent, v3 offset, v3 pos, v3* base,
                                  float* side, ray_stats* st) {
    /* the depth read and the first leaf test hit the same word: counted once */
    float extent = exp2_neg_int(nodes[parent].is_leaf_depth & DEPTH_MASK);
    offset.x = offset.x - mod1(offset.x, extent);
    offset.y = offset.y - mod1(offset.y, extent);
    offset.z = offset.z - mod1(offset.z, extent);
    for (;;) {
        st->bytes += 4;
        if (nodes[parent].is_leaf_depth >= LEAF_MASK) {
            *base = offset;
            *side = extent;
            return parent;
        }
        extent *= 0.5f;
        int mx = pos.x >= offset.x + extent;
        int my = pos.y >= offset.y + extent;
        int mz = pos.z >= offset.z + extent;
        int child = mx * 4 + my * 2 + mz;
        offset.x += (float)mx * extent;
        offset.y += (float)my * extent;
        offset.z += (float)mz * extent;
        st->bytes += 4;
        parent = nodes[parent].children[child];
    }
}

/* svo_rope.comp:50-63 */
static uint32_t neighbor_index(const uint32_t nb[3], v3 far, v3* mask) {
    if (far.x < fmin2(far.y, far.z)) {
        *mask = V3(1, 0, 0);
        return nb[0];
    } else if (far.y < far.z) {
        *mask = V3(0, 1, 0);
        return nb[1];
    } else {
        *mask = V3(0, 0, 1);
        return nb[2];
    }
}

/* svo_rope.comp:65-135 */
static v3 svo_rope_trace(const frame_ctx* c, v3 ro, v3 rd, ray_stats* st) {
    const xo_node* nodes = c->vol->nodes;
    v3 sgn = V3(sign1(rd.x), sign1(rd.y), sign1(rd.z));
    uint32_t nb[3];
    nb[0] = 1u - (uint32_t)fmax2(sgn.x, 0.0f);
    nb[1] = 3u - (uint32_t)fmax2(sgn.y, 0.0f);
    nb[2] = 5u - (uint32_t)fmax2(sgn.z, 0.0f);
    sgn = vadd(sgn, v3s(0.1f));

    v3 rrd = vdiv(v3s(1.0f), rd);
    v3 bias = vmul(rrd, ro);
    v3 box_min = V3(-bias.x, -bias.y, -bias.z);
    v3 box_max = vsub(rrd, bias);
    float t_min = max_elem(vmin(box_min, box_max));
    float t_max = min_elem(vmax(box_min, box_max));
    if (t_min > t_max) return v3s(0.0f);
    t_min = fmax2(t_min, 0.0f);

    v3 pos = vadd(ro, vscale(rd, t_min));
    v3 offset;
    float side;
    uint32_t node = svo_find(nodes, pos, &offset, &side, st);

    v3 node_min = vsub(vmul(offset, rrd), bias);
    v3 node_max = vsub(vmul(vadd(offset, v3s(side)), rrd), bias);
    v3 far = vmax(node_min, node_max);
    float u_min = max_elem(vmin(node_min, node_max));
    float u_max = min_elem(far);
    float step = u_max - fmax2(u_min, 0.0f);
    st->bytes += 4;
    v3 color = unpack_rgb(nodes[node].color);
    v3 total = vscale(color, step);
    st->steps += 1;

    v3 mask;
    uint32_t n = neighbor_index(nb, far, &mask);
    st->bytes += 4;
    node = nodes[node].children[n];
    offset = vadd(offset, vscale(vmul(mask, sgn), side));

    while (node != 0) {
        pos = vadd(ro, vscale(rd, u_max));
        node = svo_find_relative(nodes, node, offset, pos, &offset, &side, st);

        node_min = vsub(vmul(offset, rrd), bias);
        node_max = vsub(vmul(vadd(offset, v3s(side)), rrd), bias);
        far = vmax(node_min, node_max);
        u_min = max_elem(vmin(node_min, node_max));
        u_max = min_elem(far);
        step = u_max - fmax2(u_min, 0.0f);
        st->bytes += 4;
        color = unpack_rgb(nodes[node].color);
        total = vadd(total, vscale(color, step));
        st->steps += 1;

        n = neighbor_index(nb, far, &mask);
        st->bytes += 4;
        node = nodes[node].children[n];
        offset = vadd(offset, vscale(vmul(mask, sgn), side));
    }
    return total;
}

/* ------------------------------------------------------------------ */
/* main() of every shader (dda.comp:55-73, svo_naive.comp:73-89, ...)  */
/* ------------------------------------------------------------------ */

static uint32_t shade_pixel(const frame_ctx* c, int traversal, uint32_t ix, uint32_t iy, ray_stats* st) {
    int32_t pxl_x = c->out.ox + (int32_t)ix;
    int32_t pxl_y = c->out.oy + (int32_t)iy;
    float uvx = (float)(pxl_x - c->disp.ox) / (float)c->disp.w;
    float uvy = (float)(pxl_y - c->disp.oy) / (float)c->disp.h;

    v3 rd = ray(c, uvx, uvy);
    v3 color;
    switch (traversal) {
        case XO_DDA: {
            /* dda.comp:65-70; textureSize = grid dims */
            float side = max_elem(V3((float)c->vol->nx, (float)c->vol->ny, (float)c->vol->nz));
            v3 ro = vscale(c->translation, side);
            float ec = voxel_emission_coeff(c, rd) / side;
            color = vscale(dda_trace(c, ro, rd, st), ec);
            break;
        }
        case XO_SVO_NAIVE:
            color = vscale(svo_naive_trace(c, c->translation, rd, st), voxel_emission_coeff(c, rd));
            break;
        case XO_SVO_DF:
            color = vscale(svo_df_trace(c, c->translation, rd, st), voxel_emission_coeff(c, rd));
            break;
        case XO_SVO_ROPE:
            color = vscale(svo_rope_trace(c, c->translation, rd, st), voxel_emission_coeff(c, rd));
            break;
        case XO_ESVO: {
            /* esvo.comp:147-154 */
            v3 ro = vadd(c->translation, v3s(1.0f));
            float t0, t1;
            aabb_intersect(v3s(1.0f), v3s(2.0f), ro, rd, &t0, &t1);
            (void)t1;
            ro = vadd(ro, vscale(rd, fmax2(t0, 0.0f)));
            color = vscale(esvo_trace(c, ro, rd, st), voxel_emission_coeff(c, rd));
            break;
        }
        default:
            color = v3s(0.0f);
    }
    return pack_unorm8(color.x) | (pack_unorm8(color.y) << 8) | (pack_unorm8(color.z) << 16) |
           (255u << 24);
}

int xo_render(int traversal, const xo_volume* vol, const xo_params* params, const xo_camera* cam,
              const xo_rect* output, const xo_rect* display, uint32_t* rgba_out, uint32_t* steps_out,
              uint64_t* bytes_out, int threads) {
    if (!vol || !params || !cam || !output || !display || !rgba_out) return -1;
    if (traversal < 0 || traversal > XO_SVO_ROPE) return -1;
    if (traversal == XO_DDA ? vol->grid == NULL : vol->nodes == NULL) return -1;

    frame_ctx c;
    c.vol = vol;
    c.p = *params;
    c.ratio = V3(params->voxel_ratio[0], params->voxel_ratio[1], params->voxel_ratio[2]);
    c.fwd = V3(cam->forward[0], cam->forward[1], cam->forward[2]);
    c.up = V3(cam->up[0], cam->up[1], cam->up[2]);
    /* src/render/Renderer.cpp:62: translation / voxel_ratio on the host */
    c.translation = vdiv(V3(cam->translation[0], cam->translation[1], cam->translation[2]), c.ratio);
    c.out = *output;
    c.disp = *display;

#ifdef _OPENMP
    int nt = threads > 0 ? threads : omp_get_max_threads();
#else
    int nt = 1;
    (void)threads;
#endif
    (void)nt;
    const int64_t H = output->h, W = output->w;
    /* parallel over pixels in chunks of 64 (not over rows: the benchmark samples 8-row bands, which
     * would keep only two threads busy) */
    const int64_t N = W * H;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
    for (int64_t i = 0; i < N; ++i) {
        const int64_t y = i / W, x = i - y * W;
        ray_stats st = {0, 0};
        uint32_t px = shade_pixel(&c, traversal, (uint32_t)x, (uint32_t)y, &st);
        rgba_out[i] = px;
        if (steps_out) steps_out[i] = st.steps;
        if (bytes_out) bytes_out[i] = st.bytes;
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* xenodon convert: octree construction                                */
/* ------------------------------------------------------------------ */

typedef struct {
    const uint8_t* grid;
    uint64_t nx, ny, nz;
    int heuristic;
    double param;
    int dag;
    xo_node* nodes;
    uint64_t count, cap;
    xo_build_stats stats;
    /* open-addressing table for the DAG cache (HashCache, OctreeConstruction.h:19-30) */
    uint32_t* table;
    uint64_t table_cap;
} build_ctx;

static uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

/* src/model/Grid.cpp:81-137 (vol_scan) -> (avg, max_diff) */
static void vol_scan(const build_ctx* b, const uint64_t o[3], uint64_t extent, uint8_t avg[4],
                     uint8_t* max_diff) {
    uint64_t x0 = umin64(b->nx, o[0]), y0 = umin64(b->ny, o[1]), z0 = umin64(b->nz, o[2]);
    uint64_t x1 = umin64(b->nx, o[0] + extent), y1 = umin64(b->ny, o[1] + extent),
             z1 = umin64(b->nz, o[2] + extent);
    uint64_t n = (x1 - x0) * (y1 - y0) * (z1 - z0);
    memset(avg, 0, 4);
    *max_diff = 0;
    if (n == 0) return;
    uint64_t acc[4] = {0, 0, 0, 0};
    uint8_t mn[4] = {255, 255, 255, 255}, mx[4] = {0, 0, 0, 0};
    for (uint64_t z = z0; z < z1; ++z)
        for (uint64_t y = y0; y < y1; ++y) {
            const uint8_t* row = b->grid + 4 * (y * b->nx + z * b->nx * b->ny);
            for (uint64_t x = x0; x < x1; ++x)
                for (int ch = 0; ch < 4; ++ch) {
                    uint8_t v = row[4 * x + ch];
                    acc[ch] += v;
                    if (v < mn[ch]) mn[ch] = v;
                    if (v > mx[ch]) mx[ch] = v;
                }
        }
    uint8_t d = 0;
    for (int ch = 0; ch < 4; ++ch) {
        avg[ch] = (uint8_t)(acc[ch] / n);
        uint8_t dd = (uint8_t)(mx[ch] - mn[ch]);
        if (dd > d) d = dd;
    }
    *max_diff = d;
}

/* src/model/Grid.cpp:139-214 (stddev_scan) -> (avg, stddev) */
static void stddev_scan(const build_ctx* b, const uint64_t o[3], uint64_t extent, uint8_t avg[4],
                        double* stddev_out) {
    uint64_t x0 = umin64(b->nx, o[0]), y0 = umin64(b->ny, o[1]), z0 = umin64(b->nz, o[2]);
    uint64_t x1 = umin64(b->nx, o[0] + extent), y1 = umin64(b->ny, o[1] + extent),
             z1 = umin64(b->nz, o[2] + extent);
    uint64_t n = (x1 - x0) * (y1 - y0) * (z1 - z0);
    memset(avg, 0, 4);
    *stddev_out = 0;
    if (n == 0) return;
    uint64_t acc[4] = {0, 0, 0, 0};
    for (uint64_t z = z0; z < z1; ++z)
        for (uint64_t y = y0; y < y1; ++y) {
            const uint8_t* row = b->grid + 4 * (y * b->nx + z * b->nx * b->ny);
            for (uint64_t x = x0; x < x1; ++x)
                for (int ch = 0; ch < 4; ++ch) acc[ch] += row[4 * x + ch];
        }
    double nd = (double)n;
    double av[4];
    for (int ch = 0; ch < 4; ++ch) av[ch] = (double)acc[ch] / nd;
    double sd = 0;
    for (uint64_t z = z0; z < z1; ++z)
        for (uint64_t y = y0; y < y1; ++y) {
            const uint8_t* row = b->grid + 4 * (y * b->nx + z * b->nx * b->ny);
            for (uint64_t x = x0; x < x1; ++x) {
                double d0 = (double)row[4 * x + 0] - av[0];
                double d1 = (double)row[4 * x + 1] - av[1];
                double d2 = (double)row[4 * x + 2] - av[2];
                double d3 = (double)row[4 * x + 3] - av[3];
                sd += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
            }
        }
    for (int ch = 0; ch < 4; ++ch) avg[ch] = (uint8_t)av[ch];
    *stddev_out = sqrt(sd / nd);
}

static uint64_t node_hash(const xo_node* n) {
    uint64_t h = 1469598103934665603ull;
    const uint32_t* w = (const uint32_t*)n;
    for (int i = 0; i < 10; ++i) {
        h ^= w[i];
        h *= 1099511628211ull;
    }
    return h;
}

static void table_grow(build_ctx* b) {
    uint64_t ncap = b->table_cap ? b->table_cap * 2 : 1024;
    uint32_t* nt = (uint32_t*)malloc(ncap * sizeof(uint32_t));
    memset(nt, 0xFF, ncap * sizeof(uint32_t));
    for (uint64_t i = 0; i < b->table_cap; ++i) {
        uint32_t idx = b->table[i];
        if (idx == 0xFFFFFFFFu) continue;
        uint64_t s = node_hash(&b->nodes[idx]) & (ncap - 1);
        while (nt[s] != 0xFFFFFFFFu) s = (s + 1) & (ncap - 1);
        nt[s] = idx;
    }
    free(b->table);
    b->table = nt;
    b->table_cap = ncap;
}

/* OctreeBuilder::insert, OctreeConstruction.h:76-90 */
static uint32_t builder_insert(build_ctx* b, const xo_node* node, int leaf) {
    uint32_t end_index = (uint32_t)b->count;
    uint32_t actual = end_index;
    uint64_t slot = 0;
    if (b->dag) {
        if ((b->count + 1) * 2 > b->table_cap) table_grow(b);
        slot = node_hash(node) & (b->table_cap - 1);
        while (b->table[slot] != 0xFFFFFFFFu) {
            if (memcmp(&b->nodes[b->table[slot]], node, sizeof(xo_node)) == 0) {
                actual = b->table[slot];
                break;
            }
            slot = (slot + 1) & (b->table_cap - 1);
        }
    }
    int inserted = actual == end_index;
    if (inserted) {
        if (b->count == b->cap) {
            b->cap = b->cap ? b->cap * 2 : 1024;
            b->nodes = (xo_node*)realloc(b->nodes, b->cap * sizeof(xo_node));
        }
        b->nodes[b->count++] = *node;
        if (b->dag) b->table[slot] = end_index;
    }
    b->stats.total_nodes++;
    if (leaf) {
        b->stats.total_leaves++;
        if (inserted) b->stats.unique_leaves++;
    }
    return actual;
}

/* detail::construct, OctreeConstruction.h:124-194 */
static uint32_t construct(build_ctx* b, const uint64_t o[3], uint64_t extent, uint64_t depth) {
    if (depth > b->stats.depth) b->stats.depth = depth;

    const int totally_in_grid = o[0] < b->nx && o[1] < b->ny && o[2] < b->nz;
    xo_node node;
    memset(&node, 0, sizeof node);
    if (!totally_in_grid) {
        node.color = 0;
        node.is_leaf_depth = LEAF_MASK | (uint32_t)depth;
        return builder_insert(b, &node, 1);
    }
    const int partly_in_grid =
        o[0] + extent <= b->nx && o[1] + extent <= b->ny && o[2] + extent <= b->nz;

    uint8_t avg[4];
    int split;
    if (b->heuristic == XO_HEUR_STD_DEV) {
        double sd;
        stddev_scan(b, o, extent, avg, &sd);
        split = sd > b->param;
    } else {
        uint8_t md;
        vol_scan(b, o, extent, avg, &md);
        split = md > (uint8_t)b->param;
    }
    uint32_t color = (uint32_t)avg[0] | ((uint32_t)avg[1] << 8) | ((uint32_t)avg[2] << 16) |
                     ((uint32_t)avg[3] << 24);

    if ((!split && partly_in_grid) || extent == 1) {
        node.color = color;
        node.is_leaf_depth = LEAF_MASK | (uint32_t)depth;
        return builder_insert(b, &node, 1);
    }
    const uint64_t h = extent / 2;
    node.color = color;
    node.is_leaf_depth = (uint32_t)depth;
    int child = 0;
    for (int xi = 0; xi < 2; ++xi)
        for (int yi = 0; yi < 2; ++yi)
            for (int zi = 0; zi < 2; ++zi) {
                uint64_t co[3] = {o[0] + (xi ? h : 0), o[1] + (yi ? h : 0), o[2] + (zi ? h : 0)};
                node.children[child++] = construct(b, co, h, depth + 1);
            }
    return builder_insert(b, &node, 0);
}

static uint64_t ceil_2pow(uint64_t x) { /* OctreeConstruction.h:199-209 */
    --x;
    x |= x >> 1;
    x |= x >> 2;
    x |= x >> 4;
    x |= x >> 8;
    x |= x >> 16;
    x |= x >> 32;
    return ++x;
}

/* Octree::find, src/model/Octree.cpp:116-153; returns the node index (0 when out of range) */
static uint64_t octree_find(const xo_node* nodes, uint64_t dim, uint64_t px, uint64_t py, uint64_t pz,
                            uint64_t max_depth) {
    uint64_t extent = dim;
    if (px >= extent || py >= extent || pz >= extent) return 0;
    uint64_t index = 0, ox = 0, oy = 0, oz = 0;
    for (;;) {
        extent /= 2;
        if ((nodes[index].is_leaf_depth & LEAF_MASK) || extent == 0 || max_depth == 0) return index;
        uint64_t ci = 0;
        if (px >= ox + extent) { ci |= 4; ox += extent; }
        if (py >= oy + extent) { ci |= 2; oy += extent; }
        if (pz >= oz + extent) { ci |= 1; oz += extent; }
        index = nodes[index].children[ci];
        --max_depth;
    }
}

/* Octree::walk_leaves_r + generate_ropes lambda, src/model/Octree.cpp:155-201.
 * size_t arithmetic wraps exactly as in the reference (pos - extent underflows to
 * a huge value, which find() rejects as out of range -> 0). */
static void rope_walk(xo_node* nodes, uint64_t dim, uint64_t px, uint64_t py, uint64_t pz,
                      uint64_t extent, uint64_t depth, uint64_t index) {
    xo_node* node = &nodes[index];
    if (node->is_leaf_depth & LEAF_MASK) {
        node->children[0] = (uint32_t)octree_find(nodes, dim, px + extent, py, pz, depth);
        node->children[1] = (uint32_t)octree_find(nodes, dim, px - extent, py, pz, depth);
        node->children[2] = (uint32_t)octree_find(nodes, dim, px, py + extent, pz, depth);
        node->children[3] = (uint32_t)octree_find(nodes, dim, px, py - extent, pz, depth);
        node->children[4] = (uint32_t)octree_find(nodes, dim, px, py, pz + extent, depth);
        node->children[5] = (uint32_t)octree_find(nodes, dim, px, py, pz - extent, depth);
        return;
    }
    uint64_t h = extent / 2;
    int child = 0;
    for (int xi = 0; xi < 2; ++xi)
        for (int yi = 0; yi < 2; ++yi)
            for (int zi = 0; zi < 2; ++zi) {
                uint32_t ci = node->children[child++];
                rope_walk(nodes, dim, px + (xi ? h : 0), py + (yi ? h : 0), pz + (zi ? h : 0), h,
                          depth + 1, ci);
            }
}

void xo_generate_ropes(xo_node* nodes, uint64_t count, uint64_t side) {
    (void)count;
    rope_walk(nodes, side, 0, 0, 0, side, 0, 0);
}

int xo_build_octree(const uint8_t* grid, uint64_t nx, uint64_t ny, uint64_t nz, int heuristic,
                    double heuristic_param, int type, xo_node** nodes_out, uint64_t* count_out,
                    uint64_t* side_out, xo_build_stats* stats_out) {
    if (!grid || !nodes_out || !count_out || !side_out || nx == 0 || ny == 0 || nz == 0) return -1;
    build_ctx b;
    memset(&b, 0, sizeof b);
    b.grid = grid;
    b.nx = nx;
    b.ny = ny;
    b.nz = nz;
    b.heuristic = heuristic;
    b.param = heuristic_param;
    b.dag = type == XO_TYPE_DAG;

    uint64_t dim = ceil_2pow(nx);
    if (ceil_2pow(ny) > dim) dim = ceil_2pow(ny);
    if (ceil_2pow(nz) > dim) dim = ceil_2pow(nz);

    uint64_t origin[3] = {0, 0, 0};
    construct(&b, origin, dim, 0);

    /* OctreeBuilder::build, OctreeConstruction.h:92-112 */
    for (uint64_t i = 0, j = b.count - 1; i < j; ++i, --j) {
        xo_node t = b.nodes[i];
        b.nodes[i] = b.nodes[j];
        b.nodes[j] = t;
    }
    const uint32_t end = (uint32_t)b.count - 1;
    for (uint64_t i = 0; i < b.count; ++i) {
        xo_node* n = &b.nodes[i];
        if (n->is_leaf_depth & LEAF_MASK) {
            for (int k = 0; k < 8; ++k) n->children[k] = 0;
        } else {
            for (int k = 0; k < 8; ++k) n->children[k] = end - n->children[k];
        }
    }
    if (type == XO_TYPE_ROPE) xo_generate_ropes(b.nodes, b.count, dim);

    free(b.table);
    *nodes_out = b.nodes;
    *count_out = b.count;
    *side_out = dim;
    if (stats_out) *stats_out = b.stats;
    return 0;
}

void xo_free(void* p) { free(p); }
