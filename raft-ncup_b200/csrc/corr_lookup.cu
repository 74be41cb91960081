// Fused multi-scale correlation lookup (never materialises the H8*W8 x H8*W8 volume).
// Replaces CorrBlock.__call__ (core/corr.py:23-44) + bilinear_sampler (core/utils/utils.py:59-73) and the
// volume/pyramid build of CorrBlock.__init__ (core/corr.py:7-21,47-55).
//
// Math: for query pixel p with centre (cx,cy) and level l
//   G[a][c] = <f1(p), f2^l(floor(cx/2^l)-4+a, floor(cy/2^l)-4+c)> / sqrt(D)   (0 outside the level image)
//   out[l*81 + i*9 + j] = bilinear blend of G[i..i+1][j..j+1] with the fractional parts of (cx/2^l, cy/2^l)
// i (slow index) offsets x, j offsets y — the reference's transposed window (corr.py:31-37).
//
// v1 kernel, exact fp32 on CUDA cores: a CTA owns an 8x8 tile of query pixels; per level it stages the
// union bounding box of the tile's windows (16 channels at a time) in shared memory and every thread
// accumulates 25 of its pixel's 100 lattice dot products; tiles whose box does not fit fall back to
// direct global loads.
#include <cuda_fp16.h>
#include "rnc_common.cuh"

namespace rnc {

constexpr int kTile = 8;                 // 8x8 query pixels per CTA
constexpr int kPix = kTile * kTile;      // 64
constexpr int kThreads = 256;            // 4 threads per pixel
constexpr int kChunk = 16;               // channels staged per pass
constexpr int kPosStride = 20;           // floats per staged position (16 + 4 pad -> conflict-free LDS.128)
constexpr int kCap = 1024;               // staged positions
constexpr int kR = 4, kS = 9, kG = 10;   // radius, window side, lattice side
constexpr int kDots = 25;                // lattice dots per thread (4 threads x 25 = 100)

struct LookupSmem {
  float box[kCap * kPosStride];          // 80 KB
  float g[kPix][kG * kG + 1];            // lattice dot products of the tile
  float frac[kPix][2];
  int bounds[4];                         // bx0, by0, bx1, by1
};

__device__ __forceinline__ float dot16(const float (&a)[kChunk], const float* __restrict__ b) {
  const float4* b4 = reinterpret_cast<const float4*>(b);
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < kChunk / 4; ++q) {
    float4 v = b4[q];
    acc = fmaf(a[4 * q + 0], v.x, acc);
    acc = fmaf(a[4 * q + 1], v.y, acc);
    acc = fmaf(a[4 * q + 2], v.z, acc);
    acc = fmaf(a[4 * q + 3], v.w, acc);
  }
  return acc;
}

__global__ void __launch_bounds__(kThreads, 2)
corr_lookup_tile_kernel(const float* __restrict__ f1_cl, const float* __restrict__ f2_pyr,
                        const float* __restrict__ coords, int B, int D, int H, int W, int levels,
                        float* __restrict__ out, int layout, int ldo, __half* __restrict__ out_lo, int lvl_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LookupSmem& sm = *reinterpret_cast<LookupSmem*>(smem_raw);

  const int tid = threadIdx.x;
  const int p = tid >> 2, s = tid & 3;
  const int b = blockIdx.z;
  const int py = blockIdx.y * kTile + (p >> 3), px = blockIdx.x * kTile + (p & 7);
  const bool valid = py < H && px < W;
  const int P = H * W;
  const int K = levels * kS * kS;
  const float scale = 1.0f / sqrtf((float)D);   // exact for D = 256 (corr.py:55)

  float cx = 0.f, cy = 0.f;
  if (valid) {
    cx = coords[((size_t)b * 2 + 0) * P + py * W + px];
    cy = coords[((size_t)b * 2 + 1) * P + py * W + px];
  }
  // keep the float->int conversions defined for wild coordinates (window then falls fully outside -> zeros)
  cx = fminf(fmaxf(cx, -1.0e6f), 1.0e6f);
  cy = fminf(fmaxf(cy, -1.0e6f), 1.0e6f);
  const float* f1p = f1_cl + ((size_t)b * P + (valid ? py * W + px : 0)) * D;

  size_t lvl_off = 0;
  float inv = 1.0f;
  for (int l = 0; l < levels; ++l) {
    const int Hl = H >> l, Wl = W >> l;
    const float* f2l = f2_pyr + lvl_off + (size_t)b * Hl * Wl * D;
    lvl_off += (size_t)B * Hl * Wl * D;

    const float sx = cx * inv, sy = cy * inv;   // exact: inv is a power of two (corr.py:35 divides by 2**i)
    inv *= 0.5f;
    const float fx0 = floorf(sx), fy0 = floorf(sy);
    const int ix0 = (int)fx0 - kR, iy0 = (int)fy0 - kR;
    if (s == 0) { sm.frac[p][0] = sx - fx0; sm.frac[p][1] = sy - fy0; }

    // ---- union bounding box of the tile's (clipped) windows
    if (tid < 4) sm.bounds[tid] = tid < 2 ? 0x7fffffff : -0x7fffffff;
    __syncthreads();
    const int wx0 = max(ix0, 0), wx1 = min(ix0 + kG - 1, Wl - 1);
    const int wy0 = max(iy0, 0), wy1 = min(iy0 + kG - 1, Hl - 1);
    const bool nonempty = valid && wx0 <= wx1 && wy0 <= wy1;
    if (s == 0 && nonempty) {
      atomicMin(&sm.bounds[0], wx0); atomicMin(&sm.bounds[1], wy0);
      atomicMax(&sm.bounds[2], wx1); atomicMax(&sm.bounds[3], wy1);
    }
    __syncthreads();
    const int bx0 = sm.bounds[0], by0 = sm.bounds[1];
    const int bw = sm.bounds[2] - bx0 + 1, bh = sm.bounds[3] - by0 + 1;
    const bool any = sm.bounds[2] >= bx0;
    const int npos = any ? bw * bh : 0;

    // ---- per-thread lattice dots: d = s*25 + t, a = d / 10 (x), c = d % 10 (y)
    int off[kDots];
#pragma unroll
    for (int t = 0; t < kDots; ++t) {
      const int d = s * kDots + t, a = d / kG, c = d % kG;
      const int X = ix0 + a, Y = iy0 + c;
      const bool in = valid && X >= 0 && X < Wl && Y >= 0 && Y < Hl;
      off[t] = in ? ((Y - by0) * bw + (X - bx0)) : -1;
    }
    float acc[kDots];
#pragma unroll
    for (int t = 0; t < kDots; ++t) acc[t] = 0.f;

    if (npos <= kCap) {
      for (int c0 = 0; c0 < D; c0 += kChunk) {
        // stage [npos][16] channels of the box
        for (int idx = tid; idx < npos * 4; idx += kThreads) {
          const int pos = idx >> 2, q = idx & 3;
          const int by = pos / bw, bx = pos - by * bw;
          const float4 v = *reinterpret_cast<const float4*>(f2l + ((size_t)(by0 + by) * Wl + (bx0 + bx)) * D + c0 + 4 * q);
          *reinterpret_cast<float4*>(&sm.box[pos * kPosStride + 4 * q]) = v;
        }
        float a[kChunk];
#pragma unroll
        for (int q = 0; q < kChunk / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(f1p + c0 + 4 * q);
          a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < kDots; ++t)
          if (off[t] >= 0) acc[t] += dot16(a, &sm.box[off[t] * kPosStride]);
        __syncthreads();
      }
    } else {
      // incoherent tile: box does not fit -> direct global loads (correct, slow, rare)
#pragma unroll
      for (int t = 0; t < kDots; ++t) {
        if (off[t] < 0) continue;
        const int pos = off[t];
        const int by = pos / bw, bx = pos - by * bw;
        const float* f2p = f2l + ((size_t)(by0 + by) * Wl + (bx0 + bx)) * D;
        float sum = 0.f;
        for (int c0 = 0; c0 < D; c0 += kChunk) {
          float a[kChunk];
#pragma unroll
          for (int q = 0; q < kChunk / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(f1p + c0 + 4 * q);
            a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
          }
          sum += dot16(a, f2p + c0);
        }
        acc[t] = sum;
      }
    }

#pragma unroll
    for (int t = 0; t < kDots; ++t) sm.g[p][s * kDots + t] = acc[t] * scale;
    __syncthreads();

    // ---- bilinear blend + store: 64 pixels x 81 taps
    for (int o = tid; o < kPix * kS * kS; o += kThreads) {
      int q, k;
      if (layout == 0) { k = o >> 6; q = o & 63; } else { q = o / (kS * kS); k = o - q * (kS * kS); }
      const int qy = blockIdx.y * kTile + (q >> 3), qx = blockIdx.x * kTile + (q & 7);
      if (qy >= H || qx >= W) continue;
      const int i = k / kS, j = k - i * kS;
      const float ax = sm.frac[q][0], ay = sm.frac[q][1];
      const float* g = &sm.g[q][i * kG + j];
      const float v = (1.f - ax) * (1.f - ay) * g[0] + ax * (1.f - ay) * g[kG] + (1.f - ax) * ay * g[1] + ax * ay * g[kG + 1];
      if (layout == 0)
        out[(((size_t)b * K + l * kS * kS + k) * H + qy) * W + qx] = v;
      else if (layout == 1)
        out[((size_t)b * P + qy * W + qx) * ldo + l * lvl_stride + k] = v;
      else {   // layout 2: exact hi/lo split halves planes (input format of the tcgen05 convolutions), in the resident
               // channel order of the tensor-core lookup: tap (i, j) -> j*8 + i for i < 8, tap (8, j) -> 72 + j
        const size_t lvl0 = ((size_t)b * P + qy * W + qx) * ldo + l * lvl_stride;
        const size_t idx = lvl0 + (i < 8 ? j * 8 + i : 72 + j);
        const float vc = fminf(fmaxf(v, -65504.f), 65504.f);
        const __half hi = __float2half_rn(vc);
        reinterpret_cast<__half*>(out)[idx] = hi;
        out_lo[idx] = __float2half_rn(vc - __half2float(hi));
        if (k < lvl_stride - kS * kS) {          // zero the pad channels of the level (tensor-core consumers read them)
          reinterpret_cast<__half*>(out)[lvl0 + kS * kS + k] = __float2half_rn(0.f);
          out_lo[lvl0 + kS * kS + k] = __float2half_rn(0.f);
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace rnc

using namespace rnc;

static int lookup_launch(const float* f1_cl, const float* f2_pyr, const float* coords,
                         int B, int D, int H, int W, int levels, int radius,
                         float* out, int layout, int ldo, __half* out_lo, void* stream, int lvl_stride = kS * kS) {
  if (B <= 0 || H <= 0 || W <= 0 || D <= 0 || (D % kChunk) != 0) return RNC_ERR_BAD_SHAPE;
  if (levels < 1 || levels > 4 || (H >> (levels - 1)) < 1 || (W >> (levels - 1)) < 1) return RNC_ERR_BAD_SHAPE;
  if (radius != kR) return RNC_ERR_UNSUPPORTED;
  if (layout < 0 || layout > 2) return RNC_ERR_BAD_SHAPE;
  if (lvl_stride < kS * kS || (layout >= 1 && ldo < levels * lvl_stride)) return RNC_ERR_BAD_SHAPE;
  if (layout == 2 && !out_lo) return RNC_ERR_BAD_POINTER;
  if (!f1_cl || !f2_pyr || !coords || !out || !aligned16(f1_cl) || !aligned16(f2_pyr)) return RNC_ERR_BAD_POINTER;
  static unsigned long long attr_done = 0;
  if (int st = ensure_dyn_smem(corr_lookup_tile_kernel, (int)sizeof(LookupSmem), &attr_done)) return st;
  dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, B);
  corr_lookup_tile_kernel<<<grid, kThreads, sizeof(LookupSmem), as_stream(stream)>>>(
      f1_cl, f2_pyr, coords, B, D, H, W, levels, out, layout, ldo, out_lo, lvl_stride);
  return after_launch();
}

extern "C" int rnc_corr_lookup_fwd(const float* f1_cl, const float* f2_pyr, const float* coords,
                                   int B, int D, int H, int W, int levels, int radius,
                                   float* out, int layout, int ldo, void* stream) {
  if (layout != 0 && layout != 1) return RNC_ERR_BAD_SHAPE;
  return lookup_launch(f1_cl, f2_pyr, coords, B, D, H, W, levels, radius, out, layout, ldo, nullptr, stream);
}

extern "C" int rnc_corr_lookup_split_fwd(const float* f1_cl, const float* f2_pyr, const float* coords,
                                         int B, int D, int H, int W, int levels, int radius,
                                         void* out_hi, void* out_lo, int ldo, int lvl_stride, void* stream) {
  return lookup_launch(f1_cl, f2_pyr, coords, B, D, H, W, levels, radius, static_cast<float*>(out_hi), 2, ldo,
                       static_cast<__half*>(out_lo), stream, lvl_stride);
}
