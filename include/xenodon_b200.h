/*
 * xenodon_b200.h -- C ABI of the B200-native volume ray-traversal path.
 *
 * This is the drop-in boundary for the hot path of Snektron/Xenodon: everything the
 * reference does between "a volume and a camera are known on the host" and "RGBA8
 * pixels / per-frame timings are back on the host".  Each entry point names the
 * reference interface it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain C types only; no C++ or torch types cross the boundary;
 *   - every function returns 0 on success and a negative xn_status on failure; the
 *     message of the last failure on the calling thread is xn_last_error();
 *   - a context (xn_ctx) is bound to one CUDA device and owns one stream; calls on
 *     one context must come from one thread at a time (the reference is
 *     single-threaded, src/main_loop.cpp:230-258); different contexts are independent;
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     XN_ERR_CUDA.
 *
 * Reference-side binding: see INTEGRATION.md.
 */
#ifndef XENODON_B200_H
#define XENODON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XN_API __attribute__((visibility("default")))

typedef enum xn_status {
    XN_OK = 0,
    XN_ERR_INVALID = -1, /* bad argument / call order */
    XN_ERR_CUDA = -2,    /* CUDA runtime failure (no device, OOM, launch error) */
    XN_ERR_IO = -3,      /* file could not be opened / read / written */
    XN_ERR_FORMAT = -4,  /* malformed TIFF / SVO / config / camera file */
    XN_ERR_LIMIT = -5    /* exceeds a format or device limit */
} xn_status;

/* Traversal ids in the order of SHADER_OPTIONS, src/main_loop.cpp:37-43.
 * Names accepted by xn_traversal_from_name are the reference's --shader values. */
typedef enum xn_traversal {
    XN_DDA = 0,       /* resources/dda.comp        "dda"       (grid volumes) */
    XN_SVO_NAIVE = 1, /* resources/svo_naive.comp  "svo-naive" (octree volumes) */
    XN_ESVO = 2,      /* resources/esvo.comp       "esvo" */
    XN_SVO_DF = 3,    /* resources/svo_df.comp     "svo-df" */
    XN_SVO_ROPE = 4   /* resources/svo_rope.comp   "svo-rope" (needs a --rope file) */
} xn_traversal;

/* src/model/Octree.h:35-45 / resources/octree.glsl:6-10: the on-disk and host node. */
typedef struct xn_node {
    uint32_t children[8];
    uint32_t color;         /* r | g<<8 | b<<16 | a<<24 */
    uint32_t is_leaf_depth; /* bit 31 = leaf, low 31 bits = depth */
} xn_node;

/* vk::Rect2D as used by Output::region(), src/backend/Output.h:17 */
typedef struct xn_rect {
    int32_t x, y;
    uint32_t w, h;
} xn_rect;

/* RenderStats, src/render/RenderStats.h:14-27 (one frame, one or more devices) */
typedef struct xn_render_stats {
    uint64_t total_rays;
    uint64_t outputs;
    double total_render_time; /* ms, summed over outputs */
    double max_render_time;   /* ms */
    double min_render_time;   /* ms */
} xn_render_stats;

typedef struct xn_ctx xn_ctx;

/* ConstructionStats, src/model/OctreeConstruction.h:50-63 */
typedef struct xn_build_stats {
    uint64_t total_leaves, unique_leaves, total_nodes, depth;
} xn_build_stats;

XN_API const char* xn_last_error(void);
XN_API const char* xn_version(void);
XN_API int xn_traversal_from_name(const char* shader_name); /* <0 if unknown */
XN_API const char* xn_traversal_name(int traversal);

/* ---- devices: Instance::physical_devices / `xenodon sysinfo` (src/sysinfo.cpp) ---- */
XN_API int xn_device_count(int* count);
XN_API int xn_device_name(int device, char* buf, size_t cap);

/* ---- context: HeadlessOutput ctor (device, queues, fence, offscreen image),
 *      src/backend/headless/HeadlessOutput.cpp:38-63 ---- */
XN_API int xn_ctx_create(int cuda_device, xn_ctx** out);
XN_API int xn_ctx_destroy(xn_ctx* ctx);
XN_API int xn_ctx_device(const xn_ctx* ctx);

/* ---- volume upload (once per device; the volume is replicated per device) ----
 * xn_upload_grid replaces DdaRaytraceResources (src/render/DdaRaytraceAlgorithm.cpp:15-97):
 *   rgba = nx*ny*nz RGBA8 voxels, index x + y*nx + z*nx*ny (src/model/Grid.h:50-52).
 *   Sizes are 64-bit: grids beyond Vulkan's 4 GB binding limit are accepted.
 * xn_upload_svo replaces SvoRaytraceResources (src/render/SvoRaytraceAlgorithm.cpp:12-48):
 *   nodes = count 40-byte nodes exactly as stored in a .svo file, node 0 = root.
 * The caller keeps ownership of the host arrays and may free them after return. */
XN_API int xn_upload_grid(xn_ctx* ctx, const uint8_t* rgba, uint64_t nx, uint64_t ny, uint64_t nz);
XN_API int xn_upload_svo(xn_ctx* ctx, const xn_node* nodes, uint64_t count, uint64_t side);
/* Grid::load_tiff + DdaRaytraceResources in one pipelined step (src/model/Grid.cpp:27-79,
 * src/render/DdaRaytraceAlgorithm.cpp:49-96): z slices are read into page-locked staging buffers
 * by worker threads and decoded on the device (sample expansion, alpha pre-multiplication,
 * bottom-up rows) straight into the resident grid, so no host copy of the volume exists and disk
 * reads overlap the transfers.  The resident voxels equal xn_tiff_read + xn_upload_grid byte for
 * byte.  dims_out (nullable) receives nx, ny, nz; seconds_out (nullable) the wall time. */
XN_API int xn_upload_grid_tiff(xn_ctx* ctx, const char* path, uint64_t dims_out[3], double* seconds_out);
/* same, from memory already resident on ctx's device (copied device-to-device) */
XN_API int xn_upload_grid_device(xn_ctx* ctx, const void* d_rgba, uint64_t nx, uint64_t ny, uint64_t nz);
XN_API int xn_upload_svo_device(xn_ctx* ctx, const void* d_nodes40, uint64_t count, uint64_t side);
/* `xenodon convert --chan-diff n [--dag | --rope]` on the GPU, from the grid resident on ctx
 * (build_octree, src/model/OctreeConstruction.h:226-237; HashCache, :19-30; Octree::generate_ropes,
 * src/model/Octree.cpp:181-201).  The node array is byte-identical to the host builder's
 * (xn_build_octree) and so to the reference's.  type: 0 sparse, 1 dag, 2 rope.  nodes_out (nullable) receives a host copy (release with xn_free); bind != 0 also
 * makes the tree the context's resident octree, without a host round trip. */
XN_API int xn_convert_resident_grid(xn_ctx* ctx, int chan_diff, int type, int bind, xn_node** nodes_out,
                                    uint64_t* count_out, uint64_t* side_out, xn_build_stats* stats_out);
/* The same with either split heuristic (ChannelDiffHeuristic / StdDevHeuristic,
 * src/model/OctreeConstruction.h:32-48): heuristic 0 = --chan-diff (param 0..255), 1 = --std-dev
 * (param >= 0).  --std-dev is decided from exact integer sums with a bound on the rounding of the
 * reference's binary64 evaluation (src/model/Grid.cpp:139-214); when the threshold lies within that
 * bound of some cell's deviation the call fails with XN_ERR_LIMIT instead of guessing, and
 * xn_build_octree (host, replays the reference's additions) decides. */
XN_API int xn_convert_resident_grid_ex(xn_ctx* ctx, int heuristic, double param, int type, int bind,
                                       xn_node** nodes_out, uint64_t* count_out, uint64_t* side_out,
                                       xn_build_stats* stats_out);
/* deterministic synthetic volumes generated directly in device memory
 * (BASELINE.md section 4): kind 0 = "bunny-CT", 1 = "TNG gas"; bound as the grid */
XN_API int xn_synth_grid_device(xn_ctx* ctx, int kind, uint64_t nx, uint64_t ny, uint64_t nz, uint32_t seed);
/* host generator producing bit-identical voxels (host memory, multi-threaded) */
XN_API int xn_synth_grid_host(int kind, uint64_t nx, uint64_t ny, uint64_t nz, uint32_t seed, uint8_t* rgba_out);
/* copy the bound grid back to the host (nx*ny*nz*4 bytes) */
XN_API int xn_download_grid(xn_ctx* ctx, uint8_t* rgba_out, uint64_t cap_bytes);

/* Residency layout of the grid in HBM.  The reference hands its grid to the driver as an
 * "optimal tiling" 3-D image (src/render/DdaRaytraceAlgorithm.cpp:16-34); here the order is
 * explicit: x-major linear (read with plain loads); 8x8x8 bricks with Morton order inside (one
 * 32-byte sector = a 2x2x2 voxel cube); or a 3-D CUDA array read through the texture units
 * (block-linear tiling, border colour 0 = the reference's clamp-to-border sampler), where the
 * texture unit does the address arithmetic, the bounds test and the byte -> float conversion.
 * The last two keep a warp's texels in few sectors whatever the ray direction.  AUTO (default)
 * keeps volumes below 2^29 voxels linear and puts larger ones in the texture residency (measured,
 * profiles/README.md).  The mode applies to the grid resident now and to later uploads; exactly one
 * copy is resident at a time.  Strict-mode images and per-ray step counts are identical in every
 * layout; fast-mode images agree within 1/255.  xn_grid_layout reports what is resident
 * (XN_GRID_LAYOUT_AUTO = no grid) and the bytes it occupies. */
enum { XN_GRID_LAYOUT_AUTO = 0, XN_GRID_LAYOUT_LINEAR = 1, XN_GRID_LAYOUT_BRICKED = 2, XN_GRID_LAYOUT_TEXTURE = 3 };
XN_API int xn_set_grid_layout(xn_ctx* ctx, int mode);
XN_API int xn_grid_layout(const xn_ctx* ctx, int* layout_out, uint64_t* resident_bytes_out);
/* The bricked index function itself (host arithmetic, no device needed): desc_out = index-bit
 * masks of x, y, z, bit position of each axis' brick field, the top axis, total voxel slots
 * (padding included).  top = -1 lets the library choose, as uploads do.  xn_brick_indices maps
 * n (x, y, z) triples to slot indices; -1 and n_axis wrap inside the axis' own bits. */
XN_API int xn_brick_layout(uint64_t nx, uint64_t ny, uint64_t nz, int top, uint64_t desc_out[8]);
XN_API int xn_brick_indices(uint64_t nx, uint64_t ny, uint64_t nz, int top, const int32_t* xyz, uint64_t n,
                            uint64_t* index_out);

/* ---- per-output uniforms: Renderer::upload_uniform_buffers, src/render/Renderer.cpp:236-268 ----
 * output  = this device's region (HeadlessConfig device{offset,extent});
 * display = union of all regions (RenderContext::calculate_display_rect,
 *           src/render/RenderContext.cpp:43-60), used for uv so tiles are seamless. */
XN_API int xn_set_target(xn_ctx* ctx, const xn_rect* output, const xn_rect* display);
/* ShaderParameters, src/render/RenderContext.h:16-20 / src/main_loop.cpp:197-201.
 * model_dim = grid dimensions for tiff volumes, side^3 for svo volumes. */
XN_API int xn_set_params(xn_ctx* ctx, const float voxel_ratio[3], const uint32_t model_dim[3],
                         float emission_coeff);
/* Arithmetic mode of the traversal kernels.  In both modes the ray geometry (which voxels /
 * nodes a ray visits and every segment length) is computed in binary32 in the operation order
 * of the reference shaders, so per-ray step counts equal the CPU restatement exactly.
 *   XN_PRECISION_STRICT: colour accumulation also follows the shader's order: images are
 *                        bit-identical to the CPU restatement of the shaders (validation mode);
 *   XN_PRECISION_FAST  : (default) the colour sum is accumulated with fused multiply-adds on the
 *                        raw 8-bit colours and scaled once per ray; within 1/255 per channel of
 *                        STRICT on every pixel (identical on the vast majority). */
enum { XN_PRECISION_FAST = 0, XN_PRECISION_STRICT = 1 };
XN_API int xn_set_precision(xn_ctx* ctx, int mode);
/* Row interleave: render only the 16-row stripes s of the output region with
 * s % count == index.  This is the partition `count` sets of 16-row `device {}` blocks would
 * express in a headless configuration, done in one launch per device so that GPUs sharing a
 * frame get balanced work (edge bands of a frame mostly miss the volume).  Pixels of stripes
 * the context does not own are left untouched.  Default (1, 0) = the whole region.
 * xn_owned_rays returns the number of pixels the context shades per frame. */
XN_API int xn_set_interleave(xn_ctx* ctx, uint32_t count, uint32_t index);
XN_API int xn_owned_rays(const xn_ctx* ctx, uint64_t* rays);
/* Redirect the render target: pixels of this context's region are stored at
 * device_ptr[y * stride_px + x] (x, y relative to the region).  device_ptr may be
 * memory of a PEER device (NVLink): the traversal kernel then stores finished pixels
 * straight into the gathering device's frame, fusing the tile gather into the kernel.
 * NULL restores the context's own offscreen target. */
XN_API int xn_set_target_buffer(xn_ctx* ctx, void* device_ptr, size_t stride_px);

/* ---- one frame: Renderer::render, src/render/Renderer.cpp:55-103 ----
 * Asynchronous: enqueues the traversal kernel bracketed by two timing events
 * (RenderStatsCollector::pre/post_dispatch, src/render/RenderStats.cpp:46-57).
 * translation is in user units; the library divides it by voxel_ratio exactly as
 * Renderer.cpp:62 does. */
XN_API int xn_render(xn_ctx* ctx, int traversal, const float forward[3], const float up[3],
                     const float translation[3]);
/* fence wait + RenderStatsCollector::collect (HeadlessOutput::synchronize,
 * src/backend/headless/HeadlessOutput.cpp:90-93; src/render/RenderStats.cpp:59-83).
 * kernel_ms (nullable) receives the device time of the last xn_render. */
XN_API int xn_sync(xn_ctx* ctx, double* kernel_ms);
/* HeadlessOutput::download, src/backend/headless/HeadlessOutput.cpp:95-136:
 * copies the region's pixels to dst[y * stride_px + x]; stride_px 0 = region width. */
XN_API int xn_download(xn_ctx* ctx, uint32_t* dst, size_t stride_px);

/* Pipelined frame output (replaces the blocking staging copy of HeadlessOutput::download,
 * src/backend/headless/HeadlessOutput.cpp:95-136, for movie rendering): renders into one of
 * three device targets used in rotation and copies the finished region to host_dst (tight rows,
 * region w*h pixels; pinned memory from xn_host_alloc makes the copy asynchronous) on a
 * second stream, so the copy of frame i overlaps the traversal of frames i+1 and i+2.  Copies
 * complete in call order; host_dst is valid after the next xn_sync (or, for
 * xn_render_download_to, once its xn_signal_after_copy flag shows).  Not combinable with
 * xn_set_target_buffer. */
XN_API int xn_render_download_async(xn_ctx* ctx, int traversal, const float forward[3], const float up[3],
                                    const float translation[3], uint32_t* host_dst);
XN_API int xn_host_alloc(size_t bytes, void** out); /* page-locked host memory */
XN_API int xn_host_free(void* p);
/* The same pipeline for a frame that several devices (or processes, one per GPU) assemble directly
 * in host memory -- HeadlessDisplay::save's composite (src/backend/headless/HeadlessDisplay.cpp:59-76)
 * without a gathering device: host_frame is the ENCLOSING frame's pixel (region.x, region.y),
 * stride_px its row stride; only the rows this context owns (all rows, or its 16-row stripes under
 * xn_set_interleave) are written, so N contexts sharing one page-locked frame fill it over N
 * PCIe links.  xn_host_register page-locks memory the caller mapped itself (e.g. POSIX shared
 * memory mapped by every process).  xn_signal_after_copy stores `value` to *host_flag (release
 * order) once every copy enqueued so far on this context has landed: the consumer polls the flags
 * instead of synchronising with the producers. */
XN_API int xn_render_download_to(xn_ctx* ctx, int traversal, const float forward[3], const float up[3],
                                 const float translation[3], uint32_t* host_frame, size_t stride_px);
XN_API int xn_signal_after_copy(xn_ctx* ctx, volatile uint32_t* host_flag, uint32_t value);
XN_API int xn_host_register(void* p, size_t bytes);
XN_API int xn_host_unregister(void* p);

/* Device-side stopwatch over any span of calls on this context: xn_mark(ctx, 0) and
 * xn_mark(ctx, 1) record events on the context's stream; xn_mark_elapsed waits for the
 * stream and returns the device time between them (ms). */
XN_API int xn_mark(xn_ctx* ctx, int which);
XN_API int xn_mark_elapsed(xn_ctx* ctx, double* ms);
/* number of traversal-kernel launches this context has enqueued so far */
XN_API int xn_launch_count(const xn_ctx* ctx, uint64_t* count);

/* Instrumented frame (not timed): per-ray trace() loop iterations and algorithmic
 * bytes (4 B per texel fetch / node-field read as the shader source writes them).
 * steps_out / bytes_out are host arrays of region w*h elements (nullable);
 * totals_out[0] = sum of steps, totals_out[1] = sum of bytes (nullable). */
XN_API int xn_render_stats_pass(xn_ctx* ctx, int traversal, const float forward[3], const float up[3],
                                const float translation[3], uint32_t* steps_out, uint64_t* bytes_out,
                                uint64_t totals_out[2]);
/* Which voxels a DDA frame fetches (not timed; texture residency only): counts_out[0] = distinct
 * voxels fetched by the frame's rays, counts_out[1] = distinct 32-byte sectors of the x-major linear
 * volume (8 consecutive voxels) holding them -- the compulsory traffic of the frame, SURVEY 8(d)'s
 * sector-granular lower bound, to set beside the measured DRAM bytes.  use_skip_table 0 = every step
 * of dda.comp:41-50 fetches (what the shader requests); 1 = the fetches the skip table leaves. */
XN_API int xn_render_touch_pass(xn_ctx* ctx, const float forward[3], const float up[3], const float translation[3],
                                int use_skip_table, uint64_t counts_out[2]);

/* ---- multi-device frame assembly: HeadlessDisplay::save's composite,
 *      src/backend/headless/HeadlessDisplay.cpp:59-76 ----
 * Gathers every context's tile into the enclosing rectangle of all regions on
 * ctxs[0]'s device (peer-to-peer copies over NVLink when available), then copies the
 * frame to host_dst (enclosing w*h pixels; pixels outside every region are
 * 0xFF000000).  All contexts must be synchronized (xn_sync) first. */
XN_API int xn_frame_gather(xn_ctx* const* ctxs, int n, uint32_t* host_dst, xn_rect* enclosing_out);

/* Inter-process variant (one process per GPU): export ctx's frame buffer so peers can
 * store into it.  handle_out receives 64 opaque bytes (cudaIpcMemHandle_t). */
XN_API int xn_frame_buffer_create(xn_ctx* ctx, uint32_t w, uint32_t h, void** device_ptr_out,
                                  uint8_t handle_out[64]);
XN_API int xn_frame_buffer_open(xn_ctx* ctx, const uint8_t handle[64], void** device_ptr_out);
XN_API int xn_frame_buffer_close(xn_ctx* ctx, void* device_ptr);
XN_API int xn_frame_buffer_read(xn_ctx* ctx, const void* device_ptr, uint32_t w, uint32_t h,
                                uint32_t* host_dst);
/* Same, asynchronous on the context's copy stream (host_dst should be pinned, xn_host_alloc):
 * ordered after the work already enqueued on the context, overlapping later launches.  The
 * data is valid after xn_copy_sync (or the next xn_sync). */
XN_API int xn_frame_buffer_read_async(xn_ctx* ctx, const void* device_ptr, uint32_t w, uint32_t h,
                                      uint32_t* host_dst);
XN_API int xn_copy_sync(xn_ctx* ctx);

/* ---- host-side data formats (API surface of the path) ---- */

/* Grid::load_tiff, src/model/Grid.cpp:27-79 (multi-directory TIFF / BigTIFF -> RGBA8 grid,
 * TIFFReadRGBAImage conventions: bottom-up rows, alpha pre-multiplied). */
XN_API int xn_tiff_info(const char* path, uint64_t dims_out[3]);
XN_API int xn_tiff_read(const char* path, uint8_t* rgba_out, uint64_t cap_bytes);
/* How xn_upload_grid_tiff will take the file (host arithmetic only): streamable = 1 when every
 * layer is uncompressed chunky 8-bit strips of one sample layout (decoded on the device), 0 when
 * the host decoder is used (tiles, mixed layouts).  format_out (nullable) = samples per pixel,
 * photometric, has alpha, alpha is unassociated, orientation flips (bit 0 rows, bit 1 columns reversed); runs_out (nullable) = number
 * of contiguous file ranges read (adjacent strips are merged). */
XN_API int xn_tiff_stream_info(const char* path, int* streamable_out, uint32_t format_out[5], uint64_t* runs_out);
/* writer used to make inputs (uncompressed contiguous RGBA, one directory per z) */
XN_API int xn_tiff_write(const char* path, const uint8_t* rgba, uint64_t nx, uint64_t ny, uint64_t nz,
                         int bigtiff);

/* Octree::load_svo / save_svo, src/model/Octree.cpp:50-114 */
XN_API int xn_svo_info(const char* path, uint64_t* side_out, uint64_t* count_out);
XN_API int xn_svo_read(const char* path, xn_node* nodes_out, uint64_t cap_nodes);
XN_API int xn_svo_write(const char* path, const xn_node* nodes, uint64_t count, uint64_t side);

/* build_octree, src/model/OctreeConstruction.h:226-237 (`xenodon convert`).
 * heuristic 0 = --chan-diff (param 0..255), 1 = --std-dev; type 0 sparse, 1 dag, 2 rope.
 * *nodes_out is allocated by the library; release with xn_free. */
XN_API int xn_build_octree(const uint8_t* rgba, uint64_t nx, uint64_t ny, uint64_t nz, int heuristic,
                           double param, int type, xn_node** nodes_out, uint64_t* count_out,
                           uint64_t* side_out, xn_build_stats* stats_out);
XN_API void xn_free(void* p);

/* HeadlessConfig, src/backend/headless/HeadlessConfig.cpp:5-28 (`device { vkindex offset extent }`) */
typedef struct xn_headless_device {
    uint32_t vkindex;
    xn_rect region;
} xn_headless_device;
XN_API int xn_headless_config_parse(const char* text, xn_headless_device* out, int cap, int* count_out);

/* ScriptCameraController, src/camera/ScriptCameraController.cpp:4-41: 9 floats per frame
 * (forward, up, translation).  frames_out receives up to cap*9 floats. */
XN_API int xn_camera_script_parse(const char* text, float* frames_out, int cap_frames, int* count_out);

/* RenderStatsAccumulator::save, src/render/RenderStats.cpp:127-159 */
XN_API int xn_stats_write(const char* path, const xn_render_stats* frames, uint64_t n_frames,
                          double wall_seconds);

/* PNG writer for HeadlessDisplay::save (RGBA8, src/backend/headless/HeadlessDisplay.cpp:78-91) */
XN_API int xn_png_write(const char* path, const uint32_t* rgba, uint32_t w, uint32_t h);

#ifdef __cplusplus
}
#endif
#endif /* XENODON_B200_H */
