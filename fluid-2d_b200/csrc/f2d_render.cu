// f2d_render.cu -- headless renderers reading the device-resident fields directly.
//
// The reference re-uploads the host grids every frame, runs one kernel, downloads the result and
// hands it to SFML (src/density_grid_renderer.cu:38-56, src/velocity_grid_renderer.cu:55-72).  Here
// the solver's own device fields are the input; the output goes to a host buffer (image / line list)
// and no window system is involved.  Not on the hot path; the arithmetic follows the SASS of the reference
// kernels (FMUL / FMNMX / F2I.U32 for the image; FMUL, IEEE divide, FADD -- no FMA -- for the segments) and the
// output is bit-identical to theirs (tests/test_gpu_render_ref.py).
#include "f2d_kernels.cuh"

namespace f2d {

// grid_to_image_kernel (src/density_grid_renderer.cu:10-29): channel = uint8(clamp(mult * d, 0, 255)), A = 255
__global__ void k_density_to_rgba(Geom g, const float* __restrict__ d, uchar4* __restrict__ img, float mr, float mg, float mb) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.rows || j >= g.cols) return;
    const float v = __ldg(d + (size_t)i * g.pitch + j);
    uchar4 px;
    px.x = static_cast<unsigned char>(fmaxf(0.f, fminf(255.f, __fmul_rn(mr, v))));
    px.y = static_cast<unsigned char>(fmaxf(0.f, fminf(255.f, __fmul_rn(mg, v))));
    px.z = static_cast<unsigned char>(fmaxf(0.f, fminf(255.f, __fmul_rn(mb, v))));
    px.w = 255;
    img[(size_t)i * g.cols + j] = px;
}

// velocity_to_lines_kernel (src/velocity_grid_renderer.cu:8-44): one segment per cell, start == end except
// on every 8th row and column where the end follows the velocity: end += 250000 * vel / sqrtf(rows*cols).
// Output: float4 (start.x, start.y, end.x, end.y) per cell; the reference's colours are constant white.
__global__ void k_velocity_to_lines(Geom g, const float* __restrict__ u, const float* __restrict__ v, float4* __restrict__ lines,
                                    float hscale, float vscale, float norm) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.rows || j >= g.cols) return;
    const int gi = g.grow0 + i;
    float4 ln;
    ln.x = __fmul_rn((float)j, hscale);
    ln.y = __fmul_rn((float)gi, vscale);
    ln.z = ln.x;
    ln.w = ln.y;
    if (gi % 8 == 0 && j % 8 == 0) {
        const size_t o = (size_t)i * g.pitch + j;
        ln.z = __fadd_rn(ln.z, __fdiv_rn(__fmul_rn(250000.f, __ldg(u + o)), norm));
        ln.w = __fadd_rn(ln.w, __fdiv_rn(__fmul_rn(250000.f, __ldg(v + o)), norm));
    }
    lines[(size_t)i * g.cols + j] = ln;
}

void launch_density_to_rgba(const Geom& g, const float* d, void* img, float mr, float mg, float mb, cudaStream_t st) {
    dim3 bl(32, 8), gr((g.cols + 31) / 32, (g.rows + 7) / 8);
    k_density_to_rgba<<<gr, bl, 0, st>>>(g, d, static_cast<uchar4*>(img), mr, mg, mb);
}

void launch_velocity_to_lines(const Geom& g, const float* u, const float* v, void* lines, float hscale, float vscale,
                              float norm, cudaStream_t st) {
    dim3 bl(32, 8), gr((g.cols + 31) / 32, (g.rows + 7) / 8);
    k_velocity_to_lines<<<gr, bl, 0, st>>>(g, u, v, static_cast<float4*>(lines), hscale, vscale, norm);
}

}  // namespace f2d
