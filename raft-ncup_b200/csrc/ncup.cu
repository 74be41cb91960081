// Fused NCUP upsampler: zero-stuffing + the 4 live normalized-convolution layers + final scale, one kernel.
// Replaces NConvUpsampler.forward / get_out_tensor (core/upsampler.py:143-177,179-210), NConvUNet.forward
// (core/nconv_modules.py:106-136 — at the shipped config the pooled branch is dead, only
// nconv_in -> nconv_x2[0] -> decoder[0](cat(x1,x1)) -> nconv_out is live) and
// NConv2d.forward (core/nconv_modules.py:164-199):
//     den = conv(c, W), num = conv(x*c, W), y = num / (den + 1e-20), c' = den / sum(W[o])
// and the `8 *` of raft_nc_dbl.py:161.
//
// A CTA produces a 32x32 full-resolution tile of one (batch, channel) plane.  Stage 0 stages the <= 12x12
// lattice samples that can influence the tile; stage 1 (5x5 over the stride-4 lattice: <= 4 non-zero taps) fills
// a 38x38 region in shared memory; stage 2 (dense 5x5, 2->2 ch) a 34x34 region; stage 3 (3x3 with the folded
// decoder weights) and stage 4 (1x1) run per output pixel.  Positions outside the image hold zeros, which is
// exactly F.conv2d's zero padding of both the data*conf and the conf stream.
#include "rnc_common.cuh"

namespace rnc {

constexpr int NT = 32;             // output tile
constexpr int R1 = NT + 6;         // stage-1 region side (halo 3)
constexpr int R2 = NT + 2;         // stage-2 region side (halo 1)
constexpr int LT = 12;             // lattice samples per side
constexpr float kEps = 1e-20f;     // nconv_modules.py:149

struct NcupWeights {               // lives in the kernel parameter (constant) bank
  float w1[2][25];                 // nconv_in   [2,1,5,5]
  float w2[2][2][25];              // nconv_x2.0 [2,2,5,5]
  float w3[2][2][9];               // decoder.0  [2,4,3,3] folded: W[:, :2] + W[:, 2:]  (input is cat(x1, x1))
  float w4[2];                     // nconv_out  [1,2,1,1]
  float inv_s1[2], inv_s2[2], inv_s3[2];   // 1 / sum over (in,kh,kw) of the UNFOLDED weights (nconv_modules.py:186-190)
};

__global__ void __launch_bounds__(256)
ncup_fused_kernel(const float* __restrict__ x_lowres, const float* __restrict__ conf, const NcupWeights w,
                  int H4, int W4, float out_scale, float* __restrict__ out) {
  __shared__ float lx[LT][LT], lc[LT][LT];           // lattice data (flow) and confidence
  __shared__ float s1p[2][R1][R1 + 1], s1c[2][R1][R1 + 1];   // stage 1: data*conf, conf
  __shared__ float s2p[2][R2][R2 + 1], s2c[2][R2][R2 + 1];   // stage 2

  const int tid = threadIdx.x;
  const int plane = blockIdx.z;                      // b*2 + c   (channels_to_batch, upsampler.py:168)
  const int ty0 = blockIdx.y * NT, tx0 = blockIdx.x * NT;
  const int H = 4 * H4, W = 4 * W4;             // scale 4, samples at offset 2 (upsampler.py:208)
  const int iy_base = (ty0 >> 2) - 2, ix_base = (tx0 >> 2) - 2;

  // ---- stage 0: lattice samples.  X[4i+2][4j+2] = x_lowres[i][j], C[4i+2][4j+2] = conf[i][j]
  if (tid < LT * LT) {
    const int li = tid / LT, lj = tid - li * LT;
    const int iy = iy_base + li, ix = ix_base + lj;
    float xv = 0.f, cv = 0.f;
    if (iy >= 0 && iy < H4 && ix >= 0 && ix < W4) {
      xv = x_lowres[((size_t)plane * H4 + iy) * W4 + ix];
      cv = conf[((size_t)plane * H4 + iy) * W4 + ix];
    }
    lx[li][lj] = xv;
    lc[li][lj] = cv;
  }
  __syncthreads();

  // ---- stage 1: NConv(1->2, 5x5) on the zero-stuffed lattice
  for (int idx = tid; idx < R1 * R1; idx += 256) {
    const int ry = idx / R1, rx = idx - ry * R1;
    const int y = ty0 - 3 + ry, x = tx0 - 3 + rx;
    float p0 = 0.f, p1 = 0.f, c0 = 0.f, c1 = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      float n0 = 0.f, n1 = 0.f, d0 = 0.f, d1 = 0.f;
      for (int iy = max((y - 1) >> 2, 0); iy <= min(y >> 2, H4 - 1); ++iy) {
        const int ky = 4 * iy + 2 - y + 2;
        for (int ix = max((x - 1) >> 2, 0); ix <= min(x >> 2, W4 - 1); ++ix) {
          const int kx = 4 * ix + 2 - x + 2;
          const float cv = lc[iy - iy_base][ix - ix_base];
          const float xc = lx[iy - iy_base][ix - ix_base] * cv;
          const float wa = w.w1[0][ky * 5 + kx], wb = w.w1[1][ky * 5 + kx];
          d0 = fmaf(cv, wa, d0); n0 = fmaf(xc, wa, n0);
          d1 = fmaf(cv, wb, d1); n1 = fmaf(xc, wb, n1);
        }
      }
      c0 = d0 * w.inv_s1[0]; c1 = d1 * w.inv_s1[1];
      p0 = n0 / (d0 + kEps) * c0; p1 = n1 / (d1 + kEps) * c1;
    }
    s1p[0][ry][rx] = p0; s1p[1][ry][rx] = p1;
    s1c[0][ry][rx] = c0; s1c[1][ry][rx] = c1;
  }
  __syncthreads();

  // ---- stage 2: NConv(2->2, 5x5)
  for (int idx = tid; idx < R2 * R2; idx += 256) {
    const int ry = idx / R2, rx = idx - ry * R2;
    const int y = ty0 - 1 + ry, x = tx0 - 1 + rx;
    float p0 = 0.f, p1 = 0.f, c0 = 0.f, c1 = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      float n0 = 0.f, n1 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) {
            const float pv = s1p[i][ry + ky][rx + kx], cv = s1c[i][ry + ky][rx + kx];
            const float wa = w.w2[0][i][ky * 5 + kx], wb = w.w2[1][i][ky * 5 + kx];
            d0 = fmaf(cv, wa, d0); n0 = fmaf(pv, wa, n0);
            d1 = fmaf(cv, wb, d1); n1 = fmaf(pv, wb, n1);
          }
      c0 = d0 * w.inv_s2[0]; c1 = d1 * w.inv_s2[1];
      p0 = n0 / (d0 + kEps) * c0; p1 = n1 / (d1 + kEps) * c1;
    }
    s2p[0][ry][rx] = p0; s2p[1][ry][rx] = p1;
    s2c[0][ry][rx] = c0; s2c[1][ry][rx] = c1;
  }
  __syncthreads();

  // ---- stage 3 (3x3, folded decoder) + stage 4 (1x1) + scale
  for (int idx = tid; idx < NT * NT; idx += 256) {
    const int oy = idx >> 5, ox = idx & 31;
    const int y = ty0 + oy, x = tx0 + ox;
    if (y >= H || x >= W) continue;
    float n0 = 0.f, n1 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float pv = s2p[i][oy + ky][ox + kx], cv = s2c[i][oy + ky][ox + kx];
          const float wa = w.w3[0][i][ky * 3 + kx], wb = w.w3[1][i][ky * 3 + kx];
          d0 = fmaf(cv, wa, d0); n0 = fmaf(pv, wa, n0);
          d1 = fmaf(cv, wb, d1); n1 = fmaf(pv, wb, n1);
        }
    const float c0 = d0 * w.inv_s3[0], c1 = d1 * w.inv_s3[1];
    const float y0 = n0 / (d0 + kEps), y1 = n1 / (d1 + kEps);
    const float den = fmaf(c0, w.w4[0], c1 * w.w4[1]);
    const float num = fmaf(y0 * c0, w.w4[0], y1 * c1 * w.w4[1]);
    out[((size_t)plane * H + y) * W + x] = out_scale * (num / (den + kEps));
  }
}

}  // namespace rnc

using namespace rnc;

extern "C" int rnc_ncup_fwd(const float* x_lowres, const float* conf, const float* wts_host, int B, int H4, int W4,
                            float out_scale, float* out, void* stream) {
  if (B <= 0 || H4 <= 0 || W4 <= 0) return RNC_ERR_BAD_SHAPE;
  if (!x_lowres || !conf || !wts_host || !out) return RNC_ERR_BAD_POINTER;
  // wts_host: softplus'd weights in state_dict order: nconv_in[2,1,5,5], nconv_x2.0[2,2,5,5], decoder.0[2,4,3,3], nconv_out[1,2,1,1]
  NcupWeights w;
  const float* p = wts_host;
  for (int o = 0; o < 2; ++o) {
    float s = 0.f;
    for (int t = 0; t < 25; ++t) { w.w1[o][t] = p[o * 25 + t]; s += p[o * 25 + t]; }
    w.inv_s1[o] = 1.0f / s;
  }
  p += 50;
  for (int o = 0; o < 2; ++o) {
    float s = 0.f;
    for (int i = 0; i < 2; ++i)
      for (int t = 0; t < 25; ++t) { w.w2[o][i][t] = p[(o * 2 + i) * 25 + t]; s += p[(o * 2 + i) * 25 + t]; }
    w.inv_s2[o] = 1.0f / s;
  }
  p += 100;
  for (int o = 0; o < 2; ++o) {
    float s = 0.f;
    for (int i = 0; i < 4; ++i)
      for (int t = 0; t < 9; ++t) s += p[(o * 4 + i) * 9 + t];
    for (int i = 0; i < 2; ++i)
      for (int t = 0; t < 9; ++t) w.w3[o][i][t] = p[(o * 4 + i) * 9 + t] + p[(o * 4 + i + 2) * 9 + t];
    w.inv_s3[o] = 1.0f / s;
  }
  p += 72;
  w.w4[0] = p[0]; w.w4[1] = p[1];
  const int H = 4 * H4, W = 4 * W4;
  dim3 grid((W + NT - 1) / NT, (H + NT - 1) / NT, B * 2);
  ncup_fused_kernel<<<grid, 256, 0, as_stream(stream)>>>(x_lowres, conf, w, H4, W4, out_scale, out);
  return after_launch();
}
