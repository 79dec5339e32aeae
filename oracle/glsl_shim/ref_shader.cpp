// ref_shader.cpp -- host harness that runs ONE of the reference's compute shaders
// (its own text, rewritten by glsl2cpp.py, compiled against glsl.hpp) over a region.
// Compiled once per shader with -DXN_SHADER_INC=\"dda.comp.inc\" -DXN_NS=dda.
// TEST INFRASTRUCTURE ONLY (see glsl.hpp).
#include "glsl.hpp"
#include <cstddef>

#define XN_CAT2(a, b) a##b
#define XN_CAT(a, b) XN_CAT2(a, b)

namespace XN_CAT(xnref_, XN_NS) {
using namespace glsl;
// the only per-invocation built-in the shaders read
static thread_local uvec3 gl_GlobalInvocationID;
#include XN_SHADER_INC
} // namespace

namespace ns = XN_CAT(xnref_, XN_NS);

struct xnref_args {
    float forward[3], up[3], translation[3]; // translation already divided by voxel_ratio
    int32_t out_ox, out_oy;
    uint32_t out_w, out_h;
    int32_t disp_ox, disp_oy;
    uint32_t disp_w, disp_h;
    float voxel_ratio[3];
    uint32_t model_dim[3];
    float emission_coeff;
    const void* volume; // RGBA8 texels (dda) or 40-byte nodes (svo shaders)
    uint64_t nx, ny, nz;
    uint32_t* rgba_out; // out_w * out_h
};

#ifdef XN_IS_DDA
static void bind_volume(const xnref_args* a) {
    ns::model.texels = (const uint8_t*)a->volume;
    ns::model.nx = a->nx;
    ns::model.ny = a->ny;
    ns::model.nz = a->nz;
}
#else
static void bind_volume(const xnref_args* a) { ns::model.nodes = (const ns::Node*)a->volume; }
#endif

extern "C" int XN_CAT(xnref_render_, XN_NS)(const xnref_args* a) {
    using namespace glsl;
    ns::push.camera.forward = vec4(a->forward[0], a->forward[1], a->forward[2], 0.f);
    ns::push.camera.up = vec4(a->up[0], a->up[1], a->up[2], 0.f);
    ns::push.camera.translation = vec4(a->translation[0], a->translation[1], a->translation[2], 0.f);
    ns::uniforms.output_region.offset = ivec2(a->out_ox, a->out_oy);
    ns::uniforms.output_region.extent = uvec2(a->out_w, a->out_h);
    ns::uniforms.display_region.offset = ivec2(a->disp_ox, a->disp_oy);
    ns::uniforms.display_region.extent = uvec2(a->disp_w, a->disp_h);
    ns::uniforms.params.voxel_ratio = vec4(a->voxel_ratio[0], a->voxel_ratio[1], a->voxel_ratio[2], 0.f);
    ns::uniforms.params.model_dim = uvec4{a->model_dim[0], a->model_dim[1], a->model_dim[2], 0u};
    ns::uniforms.params.emission_coeff = a->emission_coeff;
    ns::render_target.pixels = a->rgba_out;
    ns::render_target.width = a->out_w;
    ns::render_target.height = a->out_h;
    bind_volume(a);

    // dispatch: ceil(extent / 8) workgroups of 8x8 (src/render/Renderer.cpp:74,89)
    const int64_t gx = ((int64_t)a->out_w - 1) / 8 + 1, gy = ((int64_t)a->out_h - 1) / 8 + 1;
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int64_t wy = 0; wy < gy; ++wy)
        for (int64_t wx = 0; wx < gx; ++wx)
            for (uint32_t ly = 0; ly < 8; ++ly)
                for (uint32_t lx = 0; lx < 8; ++lx) {
                    ns::gl_GlobalInvocationID = uvec3((uint32_t)wx * 8 + lx, (uint32_t)wy * 8 + ly, 0u);
                    ns::shader_main();
                }
    return 0;
}
