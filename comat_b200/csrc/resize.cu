// Separable table-driven resampling with fused per-channel affine: the differentiable image pre-processing of the reward path,
//   Resize((384,384), BICUBIC, antialias=True) + Normalize(mean, std)   (concept_mat_utils/caption_blip.py:33-36, :45)
// replaces aten _upsample_bicubic2d_aa(+backward) + two elementwise kernels.  The filter taps (index start, count, weights)
// are built on the host exactly as aten's _compute_indices_weights_aa does (comat_b200/image_ops.py) and passed as tables,
// so the same kernel serves forward (tables out<-in) and backward (transposed tables in<-out).
#include "common.cuh"

namespace comat {

// out[b,c,oy,ox] = scale_c * sum_ky sum_kx wy[oy][ky] wx[ox][kx] in[b,c,ys[oy]+ky, xs[ox]+kx] + shift_c
__global__ void __launch_bounds__(256) resample2d_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                         const int* __restrict__ ys, const int* __restrict__ yc, const float* __restrict__ wy,
                                                         const int* __restrict__ xs, const int* __restrict__ xc, const float* __restrict__ wx,
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         int BC, int C, int IH, int IW, int OH, int OW, int KY, int KX) {
  pdl_grid_dependency_sync();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)BC * OH * OW) return;
  const int ox = (int)(idx % OW), oy = (int)((idx / OW) % OH), bc = (int)(idx / ((long long)OW * OH));
  const float* src = in + (size_t)bc * IH * IW;
  const int y0 = ys[oy], ny = yc[oy], x0 = xs[ox], nx = xc[ox];
  float acc = 0.f;
  for (int ky = 0; ky < ny; ++ky) {
    const float* row = src + (size_t)(y0 + ky) * IW + x0;
    float r = 0.f;
    for (int kx = 0; kx < nx; ++kx) r += wx[ox * KX + kx] * row[kx];
    acc += wy[oy * KY + ky] * r;
  }
  const int c = bc % C;
  out[idx] = acc * (scale ? scale[c] : 1.f) + (shift ? shift[c] : 0.f);
}

}  // namespace comat

extern "C" int comat_resample2d(const float* in, float* out, const int* ys, const int* yc, const float* wy, const int* xs, const int* xc,
                                const float* wx, const float* scale, const float* shift, int B, int C, int IH, int IW, int OH, int OW,
                                int KY, int KX, void* stream) {
  if (!in || !out || !ys || !yc || !wy || !xs || !xc || !wx || B <= 0 || C <= 0) return COMAT_ERR_INVALID;
  const long long total = (long long)B * C * OH * OW;
  launch_k(comat::resample2d_kernel, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream, in, out, ys, yc, wy, xs, xc, wx, scale, shift,
                                                                                             B * C, C, IH, IW, OH, OW, KY, KX);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}
