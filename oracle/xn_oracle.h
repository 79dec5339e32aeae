/*
 * xn_oracle.h -- CPU restatement of Xenodon's volume ray-traversal shaders.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product
 * path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * the reported CPU baseline.
 *
 * PARITY PINNING: the reference (Snektron/Xenodon) ships no tests, golden
 * images or known-answer vectors for this path, and its
 * GLSL cannot be compiled by a GLSL compiler in this image (no glslc, no
 * Vulkan).  The restatement below is therefore pinned in two ways:
 *   1. analytically derived known answers (tests/test_oracle_known_answers.py);
 *   2. against oracle/_ref/libxnref_glsl.so, which compiles the reference's
 *      OWN shader text (resources/ *.comp, *.glsl, read where it lies under
 *      /root/reference) as C++ through the GLSL-semantics shim in
 *      oracle/glsl_shim/ (tests/test_oracle_vs_ref.py, golden fixtures in
 *      tests/golden/).
 *
 * All arithmetic is IEEE binary32, evaluated in the order the shader source
 * writes it, with no FMA contraction (-ffp-contract=off).
 */
#ifndef XN_ORACLE_H
#define XN_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* traversal ids; order of SHADER_OPTIONS in src/main_loop.cpp:37-43 */
enum {
    XO_DDA = 0,
    XO_SVO_NAIVE = 1,
    XO_ESVO = 2,
    XO_SVO_DF = 3,
    XO_SVO_ROPE = 4
};

/* src/model/Octree.h:35-45, resources/octree.glsl:6-10 */
typedef struct {
    uint32_t children[8];
    uint32_t color;         /* r | g<<8 | b<<16 | a<<24, src/model/Pixel.h:18-21 */
    uint32_t is_leaf_depth; /* bit 31 = leaf, low 31 bits = depth */
} xo_node;

/* resources/common.glsl:12-15 */
typedef struct {
    int32_t ox, oy;
    uint32_t w, h;
} xo_rect;

/* resources/common.glsl:17-21 (RenderParameters) */
typedef struct {
    float voxel_ratio[3];
    uint32_t model_dim[3];
    float emission_coeff;
} xo_params;

/* src/camera/Camera.h:6-10; translation is in user units and is divided by
 * voxel_ratio inside xo_render exactly as src/render/Renderer.cpp:62 does. */
typedef struct {
    float forward[3];
    float up[3];
    float translation[3];
} xo_camera;

typedef struct {
    const uint8_t* grid; /* RGBA8, index x + y*nx + z*nx*ny (src/model/Grid.h:50-52) */
    uint64_t nx, ny, nz;
    const xo_node* nodes; /* node 0 = root */
    uint64_t num_nodes;
} xo_volume;

/*
 * Render `output` (a sub-rectangle of the global `display` rectangle) with one
 * of the five traversals.  rgba_out: output.w*output.h packed RGBA8 pixels
 * (what imageStore to an rgba8 image leaves in memory).  steps_out (nullable):
 * per-ray loop-iteration count of trace().  bytes_out (nullable): per-ray
 * algorithmic bytes requested from the volume binding (4 B per texel fetch /
 * per node-field read as the shader source writes them).
 * threads <= 0 means "all cores".  Returns 0, or -1 on bad arguments.
 */
int xo_render(int traversal, const xo_volume* vol, const xo_params* params,
              const xo_camera* cam, const xo_rect* output, const xo_rect* display,
              uint32_t* rgba_out, uint32_t* steps_out, uint64_t* bytes_out, int threads);

/* ---- octree construction (src/model/OctreeConstruction.h, Grid.cpp:81-214) ---- */

enum { XO_TYPE_SPARSE = 0, XO_TYPE_DAG = 1, XO_TYPE_ROPE = 2 };
enum { XO_HEUR_CHAN_DIFF = 0, XO_HEUR_STD_DEV = 1 };

typedef struct {
    uint64_t total_leaves, unique_leaves, total_nodes, depth;
} xo_build_stats;

/* Builds the octree exactly as `xenodon convert` does.  On success *nodes_out
 * is malloc'ed (free with xo_free), *count_out / *side_out are set. */
int xo_build_octree(const uint8_t* grid, uint64_t nx, uint64_t ny, uint64_t nz,
                    int heuristic, double heuristic_param, int type,
                    xo_node** nodes_out, uint64_t* count_out, uint64_t* side_out,
                    xo_build_stats* stats_out);

/* src/model/Octree.cpp:181-201 on an existing tree (in place). */
void xo_generate_ropes(xo_node* nodes, uint64_t count, uint64_t side);

void xo_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
