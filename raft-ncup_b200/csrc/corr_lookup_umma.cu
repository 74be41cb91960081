// Tensor-core fused correlation lookup (v2 of corr_lookup.cu; same math; replaces
// core/corr.py:7-55 + core/utils/utils.py:59-73).  The 4-D volume is never built: per tile of 8x16 query pixels the
// kernel computes, on tcgen05, the dot products of the tile's 128 fmap1 vectors against a bounding box of fmap2^l
// positions that covers every pixel's 10x10 lattice window, then each pixel gathers its own window from the
// accumulator and blends it bilinearly.
//
//   A  = fmap1 tile      [128 px][256 ch] halves, TMA box {64,16,8,1} x 4 K-blocks, resident for the whole tile
//   B  = fmap2^l box     chunks of 256 positions x 256 ch, TMA boxes {64, BW, CR, 1}; out-of-image positions are
//                        zero-filled by TMA = the reference's zero padding (grid_sample padding_mode='zeros')
//   D  = [128][256] fp32 in TMEM, double buffered: the MMA warp fills one chunk while the epilogue drains the other
//
// Epilogue (thread = pixel = TMEM lane): one box row (BW columns) at a time is read with tcgen05.ld, parked in a
// thread-private column of shared memory (dynamic addressing), the 10 lattice values of the pixel's window row are read
// back, interpolated in x, combined with the previous row in y -> 9 outputs (fixed j, i = 0..8) into a staging tile,
// which is written out per level as coalesced hi/lo split halves (the tcgen05 convolutions' operand format).
//
// Tiles whose windows do not fit the fixed boxes (incoherent flow) are flagged and recomputed by the exact CUDA-core
// kernel (corr_lookup.cu), so the result never depends on the coherence assumption.
//
// Precision: fmaps are rounded once to fp16 (fp32 accumulate): 2.4e-4 EPE after 32 iterations in SURVEY.md Appendix D.
#include "umma_ptx.cuh"

namespace rnc {
namespace lookup_umma {

using namespace rnc::umma;

// kSplitEpi = true: 8 epilogue warps (two per TMEM lane group, splitting the slow window index i) and a 2-deep B ring;
// false: 4 epilogue warps and a 3-deep B ring.  Measured on B200 (B=8, 55x128): 0.143 ms vs 0.139 ms per launch — the
// kernel is bound by the TMA->MMA latency chain as much as by the epilogue, and shared memory cannot hold both the
// second row buffer and the third stage.
constexpr bool kSplitEpi = false;
constexpr int kEpiWarps = kSplitEpi ? 8 : 4;
constexpr int kThreads = 64 + 32 * kEpiWarps;   // warp 0 TMA, warp 1 MMA + TMEM, then the epilogue warps
constexpr int kTY = 8, kTX = 16;         // query tile (level-0 pixels)
constexpr int kD = 256;                  // feature channels
constexpr int kKB = kD / 64;             // K blocks of 64 halves
constexpr int kLevels = 4;
constexpr int kS = 9, kG = 10, kR = 4;
// per-level box: width, rows per chunk, chunks   (widths/heights seen on the benchmark stimuli: 28x24, 19x17, 15x14, 13x12)
__host__ __device__ constexpr int box_w(int l) { return l < 2 ? 32 : 16; }
__host__ __device__ constexpr int chunk_rows(int l) { return l < 2 ? 8 : 16; }
__host__ __device__ constexpr int n_chunks(int l) { return l < 2 ? 3 : 1; }
__host__ __device__ constexpr int box_h(int l) { return chunk_rows(l) * n_chunks(l); }
constexpr int kChunks = 3 + 3 + 1 + 1;
constexpr int kStages = kSplitEpi ? 2 : 3;
constexpr int kATile = 128 * 64 * 2;     // 16 KB per K block
constexpr int kBStage = 256 * 64 * 2;    // 32 KB: 256 positions x 64 halves
constexpr int kLvlStride = 88;           // channels per level in the output row (81 taps + 7 zero pads): 16-byte groups
constexpr int kSmemA = kKB * kATile;                         // 64 KB
constexpr int kSmemB = kStages * kBStage;                    // 96 KB
constexpr int kSmemScratch = (kSplitEpi ? 2 : 1) * 32 * 128 * 4;   // 16 KB row buffer per epilogue warp set
constexpr int kSmemStage = 128 * 81 * 4;                     // 41.5 KB, [tap][pixel]
constexpr int kSmemTotal = kSmemA + kSmemB + kSmemScratch + kSmemStage + 1024 + 512;

struct Params {
  const float* coords;                 // [B][2][H][W]
  __half* out_hi; __half* out_lo; int ldo;
  int* flags;                          // [tiles]: 1 = recompute this tile with the exact kernel
  int B, H, W, tiles_x, tiles_y;
  float scale;
};

struct TileInfo {                      // shared: per-level union box of the tile
  int bx0[kLevels], by0[kLevels], bx1[kLevels], by1[kLevels];
  int overflow;
};

struct EpiCtx {
  const Params& p; const TileInfo* ti; float* scratch; float* stage; uint64_t* acc_full; uint64_t* acc_empty;
  uint32_t tmem_base; int b, y0, x0, lg, ml, lane; bool valid; float cx, cy;
};

// Epilogue of one warp.  HALF selects the slow window index range this warp produces: 0 -> i in [0,5), 1 -> i in [5,9),
// 2 -> all nine (single warp per lane group).
template <int HALF>
__device__ __forceinline__ void lookup_epilogue(const EpiCtx& c) {
  constexpr int I0 = HALF == 1 ? 5 : 0, NI = HALF == 0 ? 5 : HALF == 1 ? 4 : 9;     // outputs i = I0 .. I0+NI-1 need g[I0 .. I0+NI]
  constexpr int G0 = HALF == 1 ? 6 : 0, G1 = HALF == 0 ? 6 : kLvlStride / 8;           // 8-channel output groups this warp stores
  const Params& p = c.p;
  const int HW = p.H * p.W;
  float* my_scratch = c.scratch + (HALF == 1 ? 32 * 128 : 0) + c.ml;   // [col][128] layout: thread-private column, conflict free
  float* my_stage = c.stage + c.ml;                           // [tap][128] layout; the pair shares the pixel's column
  int ch = 0;
  float inv = 1.f;
#pragma unroll 1
  for (int l = 0; l < kLevels; ++l) {
    const int Hl = p.H >> l, Wl = p.W >> l;
    const float sx = c.cx * inv, sy = c.cy * inv;
    inv *= 0.5f;
    const float fx0 = floorf(sx), fy0 = floorf(sy);
    const float ax = sx - fx0, ay = sy - fy0;
    const int ix0 = static_cast<int>(fx0) - kR, iy0 = static_cast<int>(fy0) - kR;
    const bool empty = ix0 + kG - 1 < 0 || ix0 > Wl - 1 || iy0 + kG - 1 < 0 || iy0 > Hl - 1;
    const bool any = c.ti->bx1[l] >= c.ti->bx0[l];
    const int ox = ix0 - (any ? c.ti->bx0[l] : 0), oy = iy0 - (any ? c.ti->by0[l] : 0);
    const bool live = c.valid && !empty;
    if (c.valid && empty) {
      for (int k = I0 * kS; k < (I0 + NI) * kS; ++k) my_stage[k * 128] = 0.f;
    }
    const int bw = box_w(l), cr = chunk_rows(l);
    float hprev[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) hprev[i] = 0.f;
    // warp-uniform ranges of the box rows / columns that any pixel of this warp (2 tile rows) actually needs
    const int row_lo = __reduce_min_sync(0xffffffffu, live ? oy : 0x7fffffff);
    const int row_hi = __reduce_max_sync(0xffffffffu, live ? oy + kG - 1 : -1);
    const int col_lo = __reduce_min_sync(0xffffffffu, live ? ox + I0 : 0x7fffffff);
    const int col_hi = __reduce_max_sync(0xffffffffu, live ? ox + I0 + NI : -1);
    // one box row: park the needed columns in the thread-private scratch column, gather the pixel's lattice values,
    // interpolate in x, blend with the previous row in y -> NI outputs of window row j = cidx - 1
    auto process_row = [&](const uint32_t* v, int ncols, int box_row) {
#pragma unroll
      for (int q = 0; q < 32; ++q)
        if (q < ncols && q >= col_lo && q <= col_hi) my_scratch[q * 128] = __uint_as_float(v[q]);
      const int cidx = box_row - oy;               // lattice row (y) of this pixel's window held by this box row
      if (live && cidx >= 0 && cidx < kG) {
        float g[NI + 1];
#pragma unroll
        for (int a = 0; a <= NI; ++a) g[a] = my_scratch[(ox + I0 + a) * 128];
        float h[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) h[i] = (1.f - ax) * g[i] + ax * g[i + 1];
        if (cidx >= 1) {
          const int j = cidx - 1;
#pragma unroll
          for (int i = 0; i < NI; ++i) my_stage[((I0 + i) * kS + j) * 128] = ((1.f - ay) * hprev[i] + ay * h[i]) * p.scale;
        }
#pragma unroll
        for (int i = 0; i < NI; ++i) hprev[i] = h[i];
      }
    };
    for (int cc = 0; cc < n_chunks(l); ++cc, ++ch) {
      const int buf = ch & 1, use = ch >> 1;
      mbar_wait(&c.acc_full[buf], use & 1);
      tcgen05_fence_after();
      const int r0 = max(0, row_lo - cc * cr), r1 = min(cr - 1, row_hi - cc * cr);   // rows of this chunk the warp needs
      const uint32_t tbase = c.tmem_base + (static_cast<uint32_t>(c.lg * 32) << 16) + buf * 256;
      if (r0 <= r1) {
        // software pipeline: the TMEM load of row r+1 is in flight while row r is processed
        uint32_t va[32] = {}, vb[32] = {};
        if (bw == 32) tmem_ld32_issue(tbase + r0 * 32, va); else tmem_ld16_issue(tbase + r0 * 16, va);
        tmem_ld_wait32(va);
        for (int r = r0; r <= r1; r += 2) {
          if (r + 1 <= r1) { if (bw == 32) tmem_ld32_issue(tbase + (r + 1) * 32, vb); else tmem_ld16_issue(tbase + (r + 1) * 16, vb); }
          process_row(va, bw, cc * cr + r);
          tmem_ld_wait32(vb);
          if (r + 1 <= r1) {
            if (r + 2 <= r1) { if (bw == 32) tmem_ld32_issue(tbase + (r + 2) * 32, va); else tmem_ld16_issue(tbase + (r + 2) * 16, va); }
            process_row(vb, bw, cc * cr + r + 1);
            tmem_ld_wait32(va);
          }
        }
      }
      // all of this warp's reads of the TMEM buffer are complete
      tcgen05_fence_before();
      __syncwarp();
      if (c.lane == 0) mbar_arrive(&c.acc_empty[buf]);
    }
    // ---- level done.  The level occupies kLvlStride (= 88, a multiple of 8) channels of the output row: 81 taps + 7 zero
    // pads, so every group of 8 channels is one aligned 16-byte store per plane.  The pair splits the 11 groups; a 64-thread
    // named barrier makes the partner's taps visible (and a second one protects them from the next level's writes).
    if (HALF != 2) named_bar_sync(1 + c.lg, 64);
    {
      const int q = c.ml;
      const int qy = c.y0 + (q >> 4), qx = c.x0 + (q & 15);
      if (qy < p.H && qx < p.W) {
        const size_t base = (static_cast<size_t>(c.b) * HW + qy * p.W + qx) * p.ldo + l * kLvlStride;
#pragma unroll 1
        for (int gq = G0; gq < G1; ++gq) {
          __half2 hh[4], ll[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k0 = gq * 8 + 2 * j;
            float v0 = k0 < kS * kS ? c.stage[k0 * 128 + q] : 0.f, v1 = k0 + 1 < kS * kS ? c.stage[(k0 + 1) * 128 + q] : 0.f;
            v0 = fminf(fmaxf(v0, -65504.f), 65504.f); v1 = fminf(fmaxf(v1, -65504.f), 65504.f);
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            hh[j] = __halves2half2(h0, h1);
            ll[j] = __halves2half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
          }
          *reinterpret_cast<uint4*>(p.out_hi + base + gq * 8) = *reinterpret_cast<uint4*>(hh);
          *reinterpret_cast<uint4*>(p.out_lo + base + gq * 8) = *reinterpret_cast<uint4*>(ll);
        }
      }
    }
    if (HALF != 2) named_bar_sync(1 + c.lg, 64);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
corr_lookup_umma_kernel(const __grid_constant__ CUtensorMap mF1, const __grid_constant__ CUtensorMap mL0,
                        const __grid_constant__ CUtensorMap mL1, const __grid_constant__ CUtensorMap mL2,
                        const __grid_constant__ CUtensorMap mL3, const Params p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, stays in the shared window
  unsigned char* sA = smem;
  unsigned char* sB = smem + kSmemA;
  float* scratch = reinterpret_cast<float*>(smem + kSmemA + kSmemB);
  float* stage = reinterpret_cast<float*>(smem + kSmemA + kSmemB + kSmemScratch);
  unsigned char* tail = smem + kSmemA + kSmemB + kSmemScratch + kSmemStage;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* b_full = a_full + 1;
  uint64_t* b_empty = b_full + kStages;
  uint64_t* acc_full = b_empty + kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  TileInfo* ti = reinterpret_cast<TileInfo*>(tmem_slot + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int tpi = p.tiles_x * p.tiles_y;
  const int b = tile / tpi, tr = tile - b * tpi;
  const int y0 = (tr / p.tiles_x) * kTY, x0 = (tr % p.tiles_x) * kTX;
  const int HW = p.H * p.W;

  // ---- per-pixel window origins (epilogue threads own one pixel each) and the tile's union boxes
  const bool is_epi = warp >= 2;
  const int lg = warp & 3, ml = lg * 32 + lane;
  const int py = y0 + (ml >> 4), px = x0 + (ml & 15);
  const bool valid = is_epi && py < p.H && px < p.W;
  float cx = 0.f, cy = 0.f;
  if (valid) {
    cx = p.coords[(static_cast<size_t>(b) * 2 + 0) * HW + py * p.W + px];
    cy = p.coords[(static_cast<size_t>(b) * 2 + 1) * HW + py * p.W + px];
  }
  cx = fminf(fmaxf(cx, -1.0e6f), 1.0e6f);
  cy = fminf(fmaxf(cy, -1.0e6f), 1.0e6f);

  if (threadIdx.x < kLevels) {
    ti->bx0[threadIdx.x] = 0x7fffffff; ti->by0[threadIdx.x] = 0x7fffffff;
    ti->bx1[threadIdx.x] = -0x7fffffff; ti->by1[threadIdx.x] = -0x7fffffff;
    if (threadIdx.x == 0) ti->overflow = 0;
  }
  __syncthreads();
  if (valid) {
    float inv = 1.f;
#pragma unroll
    for (int l = 0; l < kLevels; ++l) {
      const int Hl = p.H >> l, Wl = p.W >> l;
      const int ix0 = static_cast<int>(floorf(cx * inv)) - kR, iy0 = static_cast<int>(floorf(cy * inv)) - kR;
      inv *= 0.5f;
      const bool empty = ix0 + kG - 1 < 0 || ix0 > Wl - 1 || iy0 + kG - 1 < 0 || iy0 > Hl - 1;   // window fully outside: zeros
      if (!empty) {
        atomicMin(&ti->bx0[l], ix0); atomicMin(&ti->by0[l], iy0);
        atomicMax(&ti->bx1[l], ix0 + kG - 1); atomicMax(&ti->by1[l], iy0 + kG - 1);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int ov = 0;
#pragma unroll
    for (int l = 0; l < kLevels; ++l)
      if (ti->bx1[l] >= ti->bx0[l] && (ti->bx1[l] - ti->bx0[l] + 1 > box_w(l) || ti->by1[l] - ti->by0[l] + 1 > box_h(l))) ov = 1;
    ti->overflow = ov;
    p.flags[tile] = ov;
  }
  __syncthreads();
  if (ti->overflow) return;            // whole CTA: the exact kernel recomputes this tile

  if (threadIdx.x == 0) {
    mbar_init(a_full, 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(a_full, kSmemA);
#pragma unroll
      for (int kb = 0; kb < kKB; ++kb) tma_load_4d(sA + kb * kATile, &mF1, a_full, kb * 64, x0, y0, b);
      int it = 0;
#pragma unroll
      for (int l = 0; l < kLevels; ++l) {
        const CUtensorMap* map = l == 0 ? &mL0 : l == 1 ? &mL1 : l == 2 ? &mL2 : &mL3;
        const bool any = ti->bx1[l] >= ti->bx0[l];
        const int bx0 = any ? ti->bx0[l] : 0, by0 = any ? ti->by0[l] : 0;
        for (int c = 0; c < n_chunks(l); ++c)
          for (int kb = 0; kb < kKB; ++kb, ++it) {
            const int s = it % kStages, ph = (it / kStages) & 1;
            mbar_wait(&b_empty[s], ph ^ 1);
            mbar_expect_tx(&b_full[s], kBStage);
            tma_load_4d(sB + s * kBStage, map, &b_full[s], kb * 64, bx0, by0 + c * chunk_rows(l), b);
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(256 >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
      mbar_wait(a_full, 0);
      tcgen05_fence_after();
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      int it = 0;
      for (int ch = 0; ch < kChunks; ++ch) {
        const int buf = ch & 1, use = ch >> 1;
        mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
        tcgen05_fence_after();
        for (int kb = 0; kb < kKB; ++kb, ++it) {
          const int s = it % kStages, ph = (it / kStages) & 1;
          mbar_wait(&b_full[s], ph);
          tcgen05_fence_after();
          const uint64_t ad = smem_desc_sw128(a_base + kb * kATile), bd = smem_desc_sw128(b_base + s * kBStage);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_base + buf * 256, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&b_empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: gather + bilinear blend + store
    // two warps per TMEM lane group share the pixels: half 0 produces the taps i = 0..4, half 1 the taps i = 5..8
    EpiCtx c{p, ti, scratch, stage, acc_full, acc_empty, tmem_base, b, y0, x0, lg, ml, lane, valid, cx, cy};
    if (!kSplitEpi) lookup_epilogue<2>(c);
    else if (warp - 2 < 4) lookup_epilogue<0>(c);
    else lookup_epilogue<1>(c);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__global__ void f32_to_f16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, size_t n4) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = src[i];
    __half2 a = __floats2half2_rn(v.x, v.y), c = __floats2half2_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&c);
    dst[i] = o;
  }
}

}  // namespace lookup_umma
}  // namespace rnc

using namespace rnc;

extern "C" int rnc_f32_to_f16(const float* src, void* dst, size_t n, void* stream) {
  if (n == 0 || (n & 3)) return RNC_ERR_BAD_SHAPE;
  if (!src || !dst || !aligned16(src) || (reinterpret_cast<uintptr_t>(dst) & 7)) return RNC_ERR_BAD_POINTER;
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  lookup_umma::f32_to_f16_kernel<<<static_cast<int>(blocks), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(src), static_cast<uint2*>(dst), n / 4);
  return after_launch();
}

extern "C" size_t rnc_corr_lookup_umma_workspace_bytes(int B, int H, int W) {
  using namespace lookup_umma;
  return static_cast<size_t>(B) * ((H + kTY - 1) / kTY) * ((W + kTX - 1) / kTX) * sizeof(int);
}

// defined in corr_lookup.cu: exact kernel restricted to flagged 8x16 tiles, split-halves output
int rnc_corr_lookup_fallback_split(const float* f1_cl, const float* f2_pyr, const float* coords, int B, int D, int H, int W,
                                   int levels, void* out_hi, void* out_lo, int ldo, int lvl_stride, const int* flags,
                                   int flag_tiles_x, int flag_tiles_y, void* stream);

extern "C" int rnc_corr_lookup_umma_fwd(const void* f1h_cl, const void* f2h_pyr, const float* f1_cl, const float* f2_pyr,
                                        const float* coords, int B, int D, int H, int W, int levels, int radius,
                                        void* out_hi, void* out_lo, int ldo, int lvl_stride, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  using namespace lookup_umma;
  using namespace rnc::umma;
  if (B <= 0 || H <= 0 || W <= 0) return RNC_ERR_BAD_SHAPE;
  if (D != kD || levels != kLevels || radius != kR) return RNC_ERR_UNSUPPORTED;
  if ((H >> (levels - 1)) < 1 || (W >> (levels - 1)) < 1 || ldo < levels * kLvlStride || (ldo & 7) || lvl_stride != kLvlStride) return RNC_ERR_BAD_SHAPE;
  if (!f1h_cl || !f2h_pyr || !f1_cl || !f2_pyr || !coords || !out_hi || !out_lo || !workspace) return RNC_ERR_BAD_POINTER;
  if (!aligned16(f1h_cl) || !aligned16(f2h_pyr)) return RNC_ERR_BAD_POINTER;
  if (workspace_bytes < rnc_corr_lookup_umma_workspace_bytes(B, H, W)) return RNC_ERR_WORKSPACE;
  if (!encode_fn()) return RNC_ERR_UNSUPPORTED;

  Params p;
  p.coords = coords;
  p.out_hi = static_cast<__half*>(out_hi); p.out_lo = static_cast<__half*>(out_lo); p.ldo = ldo;
  p.flags = static_cast<int*>(workspace);
  p.B = B; p.H = H; p.W = W;
  p.tiles_x = (W + kTX - 1) / kTX; p.tiles_y = (H + kTY - 1) / kTY;
  p.scale = 1.0f / sqrtf(static_cast<float>(D));

  CUtensorMap maps[5];
  bool ok = make_act_map(&maps[0], f1h_cl, kD, kD, B, H, W, kTX, kTY);
  size_t off = 0;
  for (int l = 0; l < kLevels; ++l) {
    const int Hl = H >> l, Wl = W >> l;
    ok = ok && make_act_map(&maps[1 + l], static_cast<const __half*>(f2h_pyr) + off, kD, kD, B, Hl, Wl, box_w(l), chunk_rows(l));
    off += static_cast<size_t>(B) * Hl * Wl * kD;
  }
  if (!ok) return RNC_ERR_BAD_SHAPE;

  static unsigned long long done = 0;
  if (int st = ensure_dyn_smem(corr_lookup_umma_kernel, kSmemTotal, &done)) return st;
  const int ntiles = B * p.tiles_x * p.tiles_y;
  corr_lookup_umma_kernel<<<ntiles, kThreads, kSmemTotal, as_stream(stream)>>>(maps[0], maps[1], maps[2], maps[3], maps[4], p);
  if (int st = after_launch()) return st;
  // exact recomputation of the tiles the fixed boxes could not cover
  return rnc_corr_lookup_fallback_split(f1_cl, f2_pyr, coords, B, D, H, W, levels, out_hi, out_lo, ldo, lvl_stride, p.flags,
                                        p.tiles_x, p.tiles_y, stream);
}
