// Tensor-core fused correlation lookup (v2 of corr_lookup.cu; same math; replaces
// core/corr.py:7-55 + core/utils/utils.py:59-73).  The 4-D volume is never built: per tile of 8x16 query pixels the
// kernel computes, on tcgen05, the dot products of the tile's 128 fmap1 vectors against a bounding box of fmap2^l
// positions that covers every pixel's 10x10 lattice window, then each pixel gathers its own window from the
// accumulator and blends it bilinearly.
//
//   A  = fmap1 tile      [128 px][256 ch] halves, TMA box {64,16,8,1} x 4 K-blocks, resident while its levels are processed
//   B  = fmap2^l box     chunks of <= 256 positions x 256 ch streamed through a ring of 32 KB stages as 2-row TMA boxes
//                        {64, BW, 2, 1}; only the box rows some window needs are loaded and multiplied (N is set per chunk);
//                        out-of-image positions are zero-filled by TMA = the reference's zero padding (grid_sample
//                        padding_mode='zeros')
//   D  = [128][<=256] fp32 in TMEM, double buffered: the MMA warp fills one chunk while the epilogue drains the other
//
// Persistent CTAs (one per SM) walk work units (tile, level); producer / MMA / epilogue warps are decoupled by mbarrier
// rings, so the loads and MMAs of the next unit overlap the epilogue of the current one (see the kernel's comment).
//
// Epilogue (thread = pixel = TMEM lane): two box rows (BW columns each) at a time are read with tcgen05.ld and parked in
// a thread-private shared-memory row (128-bit stores; dynamic addressing is what shared memory is needed for), the 10
// lattice values of the pixel's window row are read back, interpolated in x, combined with the previous row in y -> 9
// outputs (fixed j, i = 0..8), which are split into hi/lo halves (the tcgen05 convolutions' operand format) and stored
// straight into the level's output tile in shared memory; the tile leaves with one TMA store per plane and warp.
//
// Output channel order inside a level (kLvlStride = 88 channels; consumed only by convc1, whose weights are permuted to
// match at pack time — the reference order k = i*9 + j exists only in the seam API, corr_lookup.cu):
//     tap (i, j), i < 8  ->  j*8 + i          tap (8, j)  ->  72 + j          81..87 = zero pads
// so the 8 values a thread produces per window row are ONE aligned 16-byte store per plane (conflict-free at the 176-byte
// pixel pitch) instead of 9 scattered words, and no second pass re-reads / re-packs the level.
//
// Units whose windows do not fit the fixed boxes (incoherent flow) are flagged and recomputed exactly in fp32 by the same
// CTA after its tensor-core units (exact_unit), so the result never depends on the coherence assumption.
//
// Measured on B200 (B=8, 55x128, profiles/): the three engines are balanced within 2x of each other — L2->SM operand
// traffic ~45 us, tensor pipe ~45 us, epilogue ~70 us per launch when each runs alone, ~95 us together.
//
// Precision: fmaps are rounded once to fp16 (fp32 accumulate): 2.4e-4 EPE after 32 iterations in SURVEY.md Appendix D.
#include <type_traits>

#include "umma_ptx.cuh"

namespace rnc {
namespace lookup_umma {

using namespace rnc::umma;

// RNC_LOOKUP_EPI2=1 (compile-time variant): TWO epilogue warps per TMEM lane group share every accumulator chunk — the rows a
// chunk holds for the lane group are cut in half, the halves alternate between the two warps from chunk to chunk (so each warp's
// rows are consecutive across a chunk boundary and its y-blend state carries over; the warp taking a second half re-derives the
// x-interpolated row in front of it), one box row per step with a one-row scratch each.  Bit-identical results; measured on
// B200 at B=8: 67.1 us per launch against 67.0 us for the default one-warp form (two rows per step) — after the round-2 rework
// the kernel is bound by its TMA feed (loads alone ~40 us, loads + MMA ~54 us), not by the epilogue any more, so the default
// stays the simpler form.
#ifndef RNC_LOOKUP_EPI2
#define RNC_LOOKUP_EPI2 0
#endif
constexpr int kEpiWarps = RNC_LOOKUP_EPI2 ? 8 : 4;
constexpr int kThreads = 64 + 32 * kEpiWarps;   // warp 0 TMA, warp 1 MMA + TMEM, then the epilogue warps
constexpr int kTY = 8, kTX = 16;         // query tile (level-0 pixels)
constexpr int kD = 256;                  // feature channels
constexpr int kKB = kD / 64;             // K blocks of 64 halves
constexpr int kLevels = 4;
constexpr int kS = 9, kG = 10, kR = 4;
// per-level box: width, rows per chunk, chunks   (union boxes seen on the benchmark stimuli at iteration 31: 28x24, 19x17,
// 15x14, 13x12; anything larger is handled by the exact fallback kernel)
// boxes: 32x24, 24x20, 16x16, 16x14 positions = chunks of 256, 240, 256, 224 MMA columns
#ifndef RNC_LOOKUP_CHUNKN
#define RNC_LOOKUP_CHUNKN 192    // 3 stages of 24 KB beat 2 of 32 KB (70.6 vs 75.9 us per B=8 launch): more loads in flight
#endif
__host__ __device__ constexpr int box_w(int l) { return l == 0 ? 32 : l == 1 ? 24 : 16; }
#if RNC_LOOKUP_CHUNKN == 256
__host__ __device__ constexpr int chunk_rows(int l) { return l == 0 ? 8 : l == 1 ? 10 : l == 2 ? 16 : 14; }
__host__ __device__ constexpr int n_chunks(int l) { return l == 0 ? 3 : l == 1 ? 2 : 1; }
#else   // chunks of <= 192 MMA columns (24 KB stages: a third stage fits): 6x32, 8x24, 12x16, 12x16
__host__ __device__ constexpr int chunk_rows(int l) { return l == 0 ? 6 : l == 1 ? 8 : 12; }
__host__ __device__ constexpr int n_chunks(int l) { return l == 0 ? 4 : l == 1 ? 3 : 2; }
#endif
__host__ __device__ constexpr int box_h(int l) { return chunk_rows(l) * n_chunks(l); }
__host__ __device__ constexpr int chunk_n(int l) { return box_w(l) * chunk_rows(l); }
static_assert(chunk_rows(0) <= 16 && chunk_rows(1) <= 16 && chunk_rows(2) <= 16 && chunk_rows(3) <= 16, "one box kind per even row count");
#ifndef RNC_LOOKUP_L2HINT
#define RNC_LOOKUP_L2HINT 0      // evict_last on the feature loads measured 3 % slower on B200 (81.6 vs 78.8 us per B=8 launch)
#endif
#ifndef RNC_LOOKUP_STAGES
#define RNC_LOOKUP_STAGES (RNC_LOOKUP_CHUNKN == 192 ? 3 : 2)
#endif
constexpr int kStages = RNC_LOOKUP_STAGES;   // B ring (a third stage does not fit beside the two-row scratch and measures the same)
constexpr int kATile = 128 * 64 * 2;     // 16 KB per K block
constexpr int kBStage = RNC_LOOKUP_CHUNKN * 64 * 2;    // 32 KB: 256 positions x 64 halves (24 KB with 192-column chunks)
constexpr int kLvlStride = 88;           // channels per level in the output row (81 taps + 7 zero pads): 16-byte groups
#ifndef RNC_LOOKUP_TIRING
#define RNC_LOOKUP_TIRING 4
#endif
constexpr int kTiRing = RNC_LOOKUP_TIRING;    // unit records in flight between the producer and its consumers (power of two)
constexpr int kTiShift = kTiRing == 2 ? 1 : kTiRing == 4 ? 2 : 3;
constexpr int kSmemA = kKB * kATile;                         // 64 KB
constexpr int kSmemB = kStages * kBStage;                    // 64 KB
// Epilogue buffers are pixel-major with 16-byte aligned rows, so a thread moves its data with 128-bit accesses:
//   scratch [pixel][68]: two box rows (32 words each) of the accumulator; 68 = 4*17 -> the 8 lanes of a quarter warp hit
//                        8 distinct bank quads
// scratch [pixel][68] words: two box rows (32 words each); the pixels of the odd tile row of a warp are skewed by 16 words, so
// that the gather's scalar loads (lane address = 68*pixel + skew + window offset, window offset ~ pixel's x) hit 32 distinct
// banks: bank = 5*lane + spread (mod 32)
constexpr int kScrStride = RNC_LOOKUP_EPI2 ? 36 : 68, kScrSkew = 16;   // one (EPI2) or two box rows of 32 words + 4: 36 = 68 = 4 (mod 32)
constexpr int kScrWarp = 32 * kScrStride + kScrSkew;         // words per epilogue warp (its 16 skewed pixels run 16 words past 32*68)
constexpr int kSmemScratch = kEpiWarps * kScrWarp * 4;       // 34.25 KB (4 warps x two rows) / 36.5 KB (8 warps x one row)
// output tile of one (tile, level) unit: [plane hi | lo][128 px][88 halves], dense = the TMA store's box {88, 16, 2} per warp
constexpr int kStgPlane = 128 * kLvlStride * 2;              // 22 KB
constexpr int kSmemStage = 2 * kStgPlane;                    // 44 KB
static_assert(kSmemScratch % 128 == 0 && (32 * kLvlStride * 2) % 128 == 0, "TMA store sources must be 128-byte aligned");
#ifdef RNC_PROBE_NOEPI
constexpr int kSmemTotal = kSmemA + kSmemB + 1024 + 512;     // probe: the epilogue buffers are never touched
#else
constexpr int kSmemTotal = kSmemA + kSmemB + kSmemScratch + kSmemStage + 1024 + 512;
#endif

struct Params {
  const float* f1_cl; const float* f2_pyr;   // fp32 originals: exact path of the units whose windows do not fit the boxes
  const float* coords;                 // [B][2][H][W]
  __half* out_hi; __half* out_lo; int ldo;
  int* flags;                          // [tiles][4 levels]: 1 = recompute this tile with the exact kernel
  int B, H, W, tiles_x, tiles_y;
  int tile_major;                      // unit schedule: 0 level-major, 1 tile-major, 2 hybrid (see unit_at)
  float scale;
};

// fmap2^l boxes of every even row count up to 16 per level: the rows a chunk needs are ONE TMA operation (the per-operation
// cost dominates: 2-row boxes measured 49.6 us per B=8 launch for the loads alone, whole-chunk boxes 40.9 us with more bytes)
constexpr int kBoxKinds = 8;
struct LevelMaps { CUtensorMap m[kLevels][kBoxKinds]; };

struct TileInfo {                      // shared: origin of the unit's union box (0,0 when no window is live), the box rows
  int bx0, by0, overflow, nrows;       // actually needed (even, <= box_h; 0 = no live window): only those are loaded / multiplied
};

// Window origin of one pixel at one level (shared by the producer's box computation and the epilogue's gather).
__device__ __forceinline__ bool window_origin(float cx, float cy, float inv, int Hl, int Wl, int& ix0, int& iy0) {
  ix0 = static_cast<int>(floorf(cx * inv)) - kR;
  iy0 = static_cast<int>(floorf(cy * inv)) - kR;
  return !(ix0 + kG - 1 < 0 || ix0 > Wl - 1 || iy0 + kG - 1 < 0 || iy0 > Hl - 1);   // false: window fully outside -> zeros
}

__device__ __forceinline__ float clamp_coord(float v) { return fminf(fmaxf(v, -1.0e6f), 1.0e6f); }

struct EpiCtx {
  const Params& p; const TileInfo* ti; float* scratch; __half* stage; uint64_t* acc_full; uint64_t* acc_empty;
  uint32_t tmem_base; int b, y0, x0, lg, ml, lane; bool valid; float cx, cy;
  int team;                            // EPI2: 0 / 1 = first / second epilogue warp of the lane group
};

template <int BW>
__device__ __forceinline__ void tmem_row_issue(uint32_t taddr, uint32_t* v) {
  if (BW > 16) tmem_ld32_issue(taddr, v); else tmem_ld16_issue(taddr, v);
}

// Rows of one level (box width BW columns) for the warp's 32 pixels: two box rows per step, so that two independent
// store -> gather -> interpolate chains are in flight for the single epilogue warp of each scheduler.
template <int BW>
__device__ __forceinline__ void lookup_level_rows(const EpiCtx& c, int& ch, int l, bool live, int ox, int oy, float ax, float ay) {
  const Params& p = c.p;
  const int cr = chunk_rows(l);
  float* sc = c.scratch + c.lg * kScrWarp + c.lane * kScrStride + (c.lane >> 4) * kScrSkew;   // thread-private: two box rows of 32 words
  __half* sth = c.stage + c.ml * kLvlStride;                  // this pixel's row of the level's output tile, hi plane
  __half* stl = sth + 128 * kLvlStride;                       // lo plane
  // window row j of this pixel is complete: taps (0..7, j) -> one 16-byte store per plane, tap (8, j) -> the tail
  auto emit = [&](int j, const float (&o)[kS]) {
    uint32_t hh[4], ll[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split_pair(o[2 * q], o[2 * q + 1], hh[q], ll[q]);
    *reinterpret_cast<uint4*>(sth + j * 8) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    *reinterpret_cast<uint4*>(stl + j * 8) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
    uint32_t h8, l8;
    split_pair(o[8], 0.f, h8, l8);
    reinterpret_cast<unsigned short*>(sth)[72 + j] = static_cast<unsigned short>(h8 & 0xffffu);
    reinterpret_cast<unsigned short*>(stl)[72 + j] = static_cast<unsigned short>(l8 & 0xffffu);
  };
  const float wx1 = ax * p.scale, wx0 = p.scale - wx1, wy0 = 1.f - ay;   // the 1/sqrt(D) scale rides on the x weights
  // warp-uniform range of box rows that any pixel of this warp (2 tile rows) actually needs
  const int row_lo = __reduce_min_sync(0xffffffffu, live ? oy : 0x7fffffff);
  const int row_hi = __reduce_max_sync(0xffffffffu, live ? oy + kG - 1 : -1);
  float hprev[kS];
#pragma unroll
  for (int i = 0; i < kS; ++i) hprev[i] = 0.f;
  const float* gp = sc + (live ? ox : 0);

  // one step: box rows `row`, `row + 1` (TWO) are in the scratch; cidx = lattice row of the pixel's window held by `row`
  auto step = [&](int row, auto two_tag) {
    constexpr bool TWO = decltype(two_tag)::value;
    const int cidx = row - oy;
    float g0[kG], g1[kG], h0[kS], h1[kS];
#pragma unroll
    for (int a = 0; a < kG; ++a) g0[a] = gp[a];
    if (TWO) {
#pragma unroll
      for (int a = 0; a < kG; ++a) g1[a] = gp[32 + a];
    }
#pragma unroll
    for (int i = 0; i < kS; ++i) h0[i] = wx0 * g0[i] + wx1 * g0[i + 1];
    if (TWO) {
#pragma unroll
      for (int i = 0; i < kS; ++i) h1[i] = wx0 * g1[i] + wx1 * g1[i + 1];
    }
    if (live && cidx >= 1 && cidx < kG) {
      float o[kS];
#pragma unroll
      for (int i = 0; i < kS; ++i) o[i] = wy0 * hprev[i] + ay * h0[i];
      emit(cidx - 1, o);
    }
    if (TWO && live && cidx >= 0 && cidx < kG - 1) {
      float o[kS];
#pragma unroll
      for (int i = 0; i < kS; ++i) o[i] = wy0 * h0[i] + ay * h1[i];
      emit(cidx, o);
    }
#pragma unroll
    for (int i = 0; i < kS; ++i) hprev[i] = TWO ? h1[i] : h0[i];
  };
  auto park = [&](const uint32_t* v, int half) {
#pragma unroll
    for (int g4 = 0; g4 < BW / 4; ++g4)
      *reinterpret_cast<uint4*>(sc + half * 32 + 4 * g4) = make_uint4(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3]);
  };

  const int nch = (c.ti->nrows + cr - 1) / cr;
  for (int cc = 0; cc < nch; ++cc, ++ch) {
    const int buf = ch & 1, use = ch >> 1;
    mbar_wait(&c.acc_full[buf], use & 1);
    tcgen05_fence_after();
    const int r0 = max(0, row_lo - cc * cr), r1 = min(cr - 1, row_hi - cc * cr);   // rows of this chunk the warp needs
    const uint32_t tbase = c.tmem_base + (static_cast<uint32_t>(c.lg * 32) << 16) + buf * 256;
#ifdef RNC_PROBE_NOEPI
    if (false) {
#else
    if (r0 <= r1) {
#endif
      // software pipeline: the TMEM loads of the next two rows are in flight while this pair is processed
      uint32_t va[32] = {}, vb[32] = {};
      tmem_row_issue<BW>(tbase + r0 * BW, va);
      if (r0 + 1 <= r1) tmem_row_issue<BW>(tbase + (r0 + 1) * BW, vb);
      tmem_ld_wait32(va);
      tmem_ld_wait32(vb);
      for (int r = r0; r <= r1; r += 2) {
        const bool two = r + 1 <= r1, more = r + 2 <= r1;
        park(va, 0);
        if (two) park(vb, 1);
        if (more) tmem_row_issue<BW>(tbase + (r + 2) * BW, va);
        if (r + 3 <= r1) tmem_row_issue<BW>(tbase + (r + 3) * BW, vb);
        if (two) step(cc * cr + r, std::true_type{}); else step(cc * cr + r, std::false_type{});
        if (more) { tmem_ld_wait32(va); tmem_ld_wait32(vb); }
      }
    }
    // all of this warp's reads of the TMEM buffer are complete
    tcgen05_fence_before();
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.acc_empty[buf]);
  }
}

#if RNC_LOOKUP_EPI2
// Rows of one level for one of the lane group's two warps (see RNC_LOOKUP_EPI2 above): one box row per step.
template <int BW>
__device__ __forceinline__ void lookup_level_rows2(const EpiCtx& c, int& ch, int l, bool live, int ox, int oy, float ax, float ay) {
  const Params& p = c.p;
  const int cr = chunk_rows(l);
  float* sc = c.scratch + (c.team * 4 + c.lg) * kScrWarp + c.lane * kScrStride + (c.lane >> 4) * kScrSkew;   // one box row of 32 words
  __half* sth = c.stage + c.ml * kLvlStride;
  __half* stl = sth + 128 * kLvlStride;
  auto emit = [&](int j, const float (&o)[kS]) {
    uint32_t hh[4], ll[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split_pair(o[2 * q], o[2 * q + 1], hh[q], ll[q]);
    *reinterpret_cast<uint4*>(sth + j * 8) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    *reinterpret_cast<uint4*>(stl + j * 8) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
    uint32_t h8, l8;
    split_pair(o[8], 0.f, h8, l8);
    reinterpret_cast<unsigned short*>(sth)[72 + j] = static_cast<unsigned short>(h8 & 0xffffu);
    reinterpret_cast<unsigned short*>(stl)[72 + j] = static_cast<unsigned short>(l8 & 0xffffu);
  };
  const float wx1 = ax * p.scale, wx0 = p.scale - wx1, wy0 = 1.f - ay;
  const int row_lo = __reduce_min_sync(0xffffffffu, live ? oy : 0x7fffffff);
  const int row_hi = __reduce_max_sync(0xffffffffu, live ? oy + kG - 1 : -1);
  float hprev[kS];
#pragma unroll
  for (int i = 0; i < kS; ++i) hprev[i] = 0.f;
  const float* gp = sc + (live ? ox : 0);
  // one box row: x-interpolate; unless it only primes the y-blend state, blend with the previous row and store window row cidx-1
  auto step = [&](const uint32_t* v, int row, bool prime) {
#pragma unroll
    for (int g4 = 0; g4 < BW / 4; ++g4)
      *reinterpret_cast<uint4*>(sc + 4 * g4) = make_uint4(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3]);
    const int cidx = row - oy;
    float g[kG], h[kS];
#pragma unroll
    for (int a = 0; a < kG; ++a) g[a] = gp[a];
#pragma unroll
    for (int i = 0; i < kS; ++i) h[i] = wx0 * g[i] + wx1 * g[i + 1];
    if (!prime && live && cidx >= 1 && cidx < kG) {
      float o[kS];
#pragma unroll
      for (int i = 0; i < kS; ++i) o[i] = wy0 * hprev[i] + ay * h[i];
      emit(cidx - 1, o);
    }
#pragma unroll
    for (int i = 0; i < kS; ++i) hprev[i] = h[i];
  };

  const int nch = (c.ti->nrows + cr - 1) / cr;
  for (int cc = 0; cc < nch; ++cc, ++ch) {
    const int buf = ch & 1, use = ch >> 1;
    mbar_wait(&c.acc_full[buf], use & 1);
    tcgen05_fence_after();
    const int r0 = max(0, row_lo - cc * cr), r1 = min(cr - 1, row_hi - cc * cr);   // rows of this chunk the lane group needs
    const uint32_t tbase = c.tmem_base + (static_cast<uint32_t>(c.lg * 32) << 16) + buf * 256;
    if (r0 <= r1) {
      // halves alternate between the two warps: role 0 = first half (continues this warp's rows of the previous chunk),
      // role 1 = second half (re-derives the row in front of it to prime hprev)
      const int role = (cc & 1) ^ c.team, nfirst = (r1 - r0 + 2) >> 1;
      const int a = role == 0 ? r0 : r0 + nfirst - 1, b = role == 0 ? r0 + nfirst - 1 : r1;   // role 1 starts on its priming row
      uint32_t va[32] = {}, vb[32] = {};
      tmem_row_issue<BW>(tbase + a * BW, va);
      for (int r = a; r <= b; r += 2) {
        tmem_ld_wait32(va);
        if (r + 1 <= b) tmem_row_issue<BW>(tbase + (r + 1) * BW, vb);
        step(va, cc * cr + r, role == 1 && r == a);
        if (r + 1 <= b) {
          tmem_ld_wait32(vb);
          if (r + 2 <= b) tmem_row_issue<BW>(tbase + (r + 2) * BW, va);
          step(vb, cc * cr + r + 1, false);
        }
      }
    }
    tcgen05_fence_before();
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.acc_empty[buf]);
  }
}
#endif

// Epilogue of one warp (thread = pixel = TMEM lane) for one unit = (tile, level l): rows into the output tile, then one TMA
// store per plane of the warp's 2 x 16 pixels (the TMA unit clips pixels beyond the image).
__device__ __forceinline__ void lookup_epilogue(const EpiCtx& c, int& ch, int l, const CUtensorMap* map_hi, const CUtensorMap* map_lo) {
  const Params& p = c.p;
  __half* sth = c.stage + c.ml * kLvlStride;
  __half* stl = sth + 128 * kLvlStride;
  const float inv = 1.f / static_cast<float>(1 << l);
  const float sx = c.cx * inv, sy = c.cy * inv;
  const float ax = sx - floorf(sx), ay = sy - floorf(sy);
  int ix0, iy0;
  const bool empty = !window_origin(c.cx, c.cy, inv, p.H >> l, p.W >> l, ix0, iy0);
  const int ox = ix0 - c.ti->bx0, oy = iy0 - c.ti->by0;
  const bool live = c.valid && !empty;
  // the previous unit's TMA stores have finished READING this lane group's rows of the output tile
#if RNC_LOOKUP_EPI2
  if (c.team == 0 && c.lane == 0) bulk_wait_read0();
  named_bar_sync(1 + c.lg, 64);                // both warps of the lane group: the tile may be rewritten
#else
  if (c.lane == 0) bulk_wait_read0();
  __syncwarp();
#endif
#ifndef RNC_PROBE_NOEPI
  if (!live && (!RNC_LOOKUP_EPI2 || c.team == 0)) {   // window fully outside the level image (or pixel outside the frame): zeros
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int q = 0; q < 10; ++q) {
      *reinterpret_cast<uint4*>(sth + q * 8) = z;
      *reinterpret_cast<uint4*>(stl + q * 8) = z;
    }
    reinterpret_cast<unsigned short*>(sth)[80] = 0;
    reinterpret_cast<unsigned short*>(stl)[80] = 0;
  }
#endif
#if RNC_LOOKUP_EPI2
  if (l == 0) lookup_level_rows2<box_w(0)>(c, ch, l, live, ox, oy, ax, ay);
  else if (l == 1) lookup_level_rows2<box_w(1)>(c, ch, l, live, ox, oy, ax, ay);
  else lookup_level_rows2<box_w(2)>(c, ch, l, live, ox, oy, ax, ay);
#else
  if (l == 0) lookup_level_rows<box_w(0)>(c, ch, l, live, ox, oy, ax, ay);
  else if (l == 1) lookup_level_rows<box_w(1)>(c, ch, l, live, ox, oy, ax, ay);
  else lookup_level_rows<box_w(2)>(c, ch, l, live, ox, oy, ax, ay);
#endif
  static_assert(box_w(2) == box_w(3), "levels 2 and 3 share the row code");
#ifndef RNC_PROBE_NOEPI
  fence_proxy_async();                         // this thread's shared-memory writes are visible to the TMA unit
#if RNC_LOOKUP_EPI2
  named_bar_sync(1 + c.lg, 64);                // both warps' window rows are in the tile
  if (c.team == 0 && c.lane == 0) {
#else
  __syncwarp();
  if (c.lane == 0) {
#endif
    const __half* src = c.stage + c.lg * 32 * kLvlStride;
    tma_store_4d(map_hi, src, l * kLvlStride, c.x0, c.y0 + 2 * c.lg, c.b);
    tma_store_4d(map_lo, src + 128 * kLvlStride, l * kLvlStride, c.x0, c.y0 + 2 * c.lg, c.b);
    bulk_commit();
  }
#endif
}

// Work units.  A unit is (tile, level): 3, 2, 1, 1 accumulator chunks.  Schedules (Params.tile_major):
//   0  level-major: units listed level-major (all level-0 units first = longest first) and dealt to the CTAs in snake order
//      (round k runs left-to-right for even k, right-to-left for odd k); every unit loads its own A tile
//   1  tile-major: the four levels of a tile back to back on one CTA, A loaded once per tile (4x less A traffic, no A reload
//      bubble between levels), but whole tiles balance badly (B=8: 448 tiles on 148 CTAs -> the busiest CTA gets 4)
//   2  hybrid (default): the complete rounds of tiles (floor(ntiles / CTAs) per CTA) tile-major, the remaining tiles
//      level-major over all CTAs.  B=8: 3 tile-major rounds + 16 units -> busiest CTA 24 chunks (level-major: 23) with a
//      quarter of the A loads; a single image (56 tiles < 148 CTAs) is all level-major and still fills the SMs.
// Returns the k-th unit of this CTA as (tile-major flag << 30) | (level * ntiles + tile), or -1 (none in round k).
constexpr int kUnitTm = 1 << 30;
__device__ __forceinline__ int unit_id(int u) { return u & (kUnitTm - 1); }
__device__ __forceinline__ bool unit_tm(int u) { return (u & kUnitTm) != 0; }
__device__ __forceinline__ int unit_at(int k, int ntiles, int mode) {
  const int G = gridDim.x, c = blockIdx.x;
  const int full = mode == 0 ? 0 : mode == 1 ? (ntiles + G - 1) / G : ntiles / G;      // tile-major rounds
  if (k < full * kLevels) {
    const int tile = (k / kLevels) * G + c;
    return tile < ntiles ? kUnitTm | ((k % kLevels) * ntiles + tile) : -1;
  }
  if (mode == 1) return -1;
  const int kk = k - full * kLevels, done = full * G, rem = ntiles - done;             // remaining tiles, level-major snake
  const int g = kk * G + ((kk & 1) ? G - 1 - c : c);
  if (g >= rem * kLevels) return -1;
  const int l = g / rem;
  return l * ntiles + done + (g - l * rem);
}
__device__ __forceinline__ int unit_rounds(int ntiles, int mode) {
  const int G = gridDim.x;
  if (mode == 1) return ((ntiles + G - 1) / G) * kLevels;
  const int full = mode == 0 ? 0 : ntiles / G;
  return full * kLevels + ((ntiles - full * G) * kLevels + G - 1) / G;
}

// Exact fp32 recomputation of one (tile, level) unit whose windows did not fit the fixed box (incoherent flow, e.g. a motion
// boundary): warp = pixel, the 100 lattice dot products of the pixel's window straight from the fp32 feature maps (lane = 8
// channels, positions outside the level image contribute 0 = grid_sample's zero padding), then the bilinear blend of
// SURVEY.md Appendix A.1 and the same hi/lo split output in the resident channel order.  Runs on all warps of the CTA after
// its tensor-core units: results never depend on a coherence assumption, and there is no second kernel launch.
__device__ __noinline__ void exact_unit(const Params& p, int l, int tile, float* gsm, int warp, int lane, int nwarps) {
  const int tpi = p.tiles_x * p.tiles_y, HW = p.H * p.W;
  const int b = tile / tpi, tr = tile - b * tpi;
  const int y0 = (tr / p.tiles_x) * kTY, x0 = (tr % p.tiles_x) * kTX;
  const int Hl = p.H >> l, Wl = p.W >> l;
  size_t off = 0;
  for (int k = 0; k < l; ++k) off += static_cast<size_t>(p.B) * (p.H >> k) * (p.W >> k) * kD;
  const float* f2l = p.f2_pyr + off + static_cast<size_t>(b) * Hl * Wl * kD;
  const float inv = 1.f / static_cast<float>(1 << l);
  float* g = gsm + warp * 104;
  for (int px = warp; px < kTY * kTX; px += nwarps) {
    const int qy = y0 + (px >> 4), qx = x0 + (px & 15);
    if (qy >= p.H || qx >= p.W) continue;                       // warp-uniform
    const float cx = clamp_coord(__ldg(p.coords + (static_cast<size_t>(b) * 2 + 0) * HW + qy * p.W + qx));
    const float cy = clamp_coord(__ldg(p.coords + (static_cast<size_t>(b) * 2 + 1) * HW + qy * p.W + qx));
    const float sx = cx * inv, sy = cy * inv;
    const float fx0 = floorf(sx), fy0 = floorf(sy), ax = sx - fx0, ay = sy - fy0;
    const int ix0 = static_cast<int>(fx0) - kR, iy0 = static_cast<int>(fy0) - kR;
    const float4* f1p = reinterpret_cast<const float4*>(p.f1_cl + (static_cast<size_t>(b) * HW + qy * p.W + qx) * kD);
    const float4 a0 = __ldg(f1p + lane), a1 = __ldg(f1p + 32 + lane);
    // lattice (a = x offset, c = y offset) -> g[a*10 + c]; five positions per pass so that their loads are in flight together
    for (int c = 0; c < kG; ++c) {
      const int Y = iy0 + c;
      const bool rowin = Y >= 0 && Y < Hl;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float4 b0[5], b1[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          const int X = ix0 + half * 5 + u;
          const bool in = rowin && X >= 0 && X < Wl;              // warp-uniform
          const float4* q = reinterpret_cast<const float4*>(f2l + (static_cast<size_t>(in ? Y : 0) * Wl + (in ? X : 0)) * kD);
          b0[u] = in ? __ldg(q + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
          b1[u] = in ? __ldg(q + 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          float d = a0.x * b0[u].x;
          d = fmaf(a0.y, b0[u].y, d); d = fmaf(a0.z, b0[u].z, d); d = fmaf(a0.w, b0[u].w, d);
          d = fmaf(a1.x, b1[u].x, d); d = fmaf(a1.y, b1[u].y, d); d = fmaf(a1.z, b1[u].z, d); d = fmaf(a1.w, b1[u].w, d);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
          if (lane == 0) g[(half * 5 + u) * kG + c] = d * p.scale;
        }
      }
    }
    __syncwarp();
    const size_t base = (static_cast<size_t>(b) * HW + qy * p.W + qx) * p.ldo + l * kLvlStride;
    for (int t = lane; t < kS * kS; t += 32) {
      const int i = t / kS, j = t - i * kS;
      const float* gp = g + i * kG + j;
      const float v = (1.f - ax) * (1.f - ay) * gp[0] + ax * (1.f - ay) * gp[kG] + (1.f - ax) * ay * gp[1] + ax * ay * gp[kG + 1];
      uint32_t hh, ll;
      split_pair(v, 0.f, hh, ll);
      const size_t k = base + (i < 8 ? j * 8 + i : 72 + j);
      reinterpret_cast<unsigned short*>(p.out_hi)[k] = static_cast<unsigned short>(hh & 0xffffu);
      reinterpret_cast<unsigned short*>(p.out_lo)[k] = static_cast<unsigned short>(ll & 0xffffu);
    }
    if (lane < kLvlStride - kS * kS) {                          // the level's zero pads
      reinterpret_cast<unsigned short*>(p.out_hi)[base + kS * kS + lane] = 0;
      reinterpret_cast<unsigned short*>(p.out_lo)[base + kS * kS + lane] = 0;
    }
    __syncwarp();
  }
}

// Persistent kernel: one CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The three roles run decoupled
// through mbarrier rings, so the TMA loads and MMAs of tile n+1 overlap the gather/blend/store epilogue of tile n:
//   warp 0      producer: loads the tile's coords, reduces the per-level union boxes, publishes them (ti ring, 2 slots),
//               then streams A (per K block, as soon as the previous tile's last chunk released it) and the B chunks
//   warp 1      MMA issuer; TMEM accumulators are double buffered across chunks and tiles
//   warps 2..   epilogue
__global__ void __launch_bounds__(kThreads, 1)
corr_lookup_umma_kernel(const __grid_constant__ CUtensorMap mF1, const __grid_constant__ LevelMaps mLv,
                        const __grid_constant__ CUtensorMap mOutHi,
                        const __grid_constant__ CUtensorMap mOutLo, const Params p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, stays in the shared window
  unsigned char* sA = smem;
  unsigned char* sB = smem + kSmemA;
  float* scratch = reinterpret_cast<float*>(smem + kSmemA + kSmemB);
  __half* stage = reinterpret_cast<__half*>(smem + kSmemA + kSmemB + kSmemScratch);
  unsigned char* tail = smem + kSmemTotal - 1024 - 512;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);       // [kKB]
  uint64_t* a_empty = a_full + kKB;                           // [kKB]
  uint64_t* b_full = a_empty + kKB;
  uint64_t* b_empty = b_full + kStages;
  uint64_t* acc_full = b_empty + kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* ti_full = acc_empty + 2;                          // [kTiRing]
  uint64_t* ti_empty = ti_full + kTiRing;                     // [kTiRing]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ti_empty + kTiRing);
  int* ov_count = reinterpret_cast<int*>(tmem_slot + 1);      // units of this CTA that overflowed their box
  TileInfo* ti = reinterpret_cast<TileInfo*>(tmem_slot + 2);  // [kTiRing]

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tpi = p.tiles_x * p.tiles_y;
  const int ntiles = p.B * tpi;
  const int HW = p.H * p.W;

  if (threadIdx.x == 0) {
    *ov_count = 0;
    for (int kb = 0; kb < kKB; ++kb) { mbar_init(&a_full[kb], 1); mbar_init(&a_empty[kb], 1); }
    for (int s = 0; s < kStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiWarps); }
    for (int i = 0; i < kTiRing; ++i) { mbar_init(&ti_full[i], 1); mbar_init(&ti_empty[i], kEpiWarps + 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                          // coords1 of the previous iteration's flow update is visible from here

  const int sched = p.tile_major;
  const int rounds = unit_rounds(ntiles, sched);

  if (warp == 0) {
    // ------------------------------------------------------------------ producer (warp-uniform loops, one elected lane issues)
    // lane owns pixels lane, lane+32, lane+64, lane+96 of the tile for the box reduction
    float pcx[4], pcy[4];
#if RNC_LOOKUP_L2HINT
    const uint64_t pol_keep = l2_policy_evict_last();   // the feature maps are re-read by every tile and every iteration
#endif
    auto load_coords = [&](int tile, float (&ox)[4], float (&oy)[4]) {
      const int b = tile / tpi, tr = tile - b * tpi;
      const int y0 = (tr / p.tiles_x) * kTY, x0 = (tr % p.tiles_x) * kTX;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ml = lane + 32 * j, py = y0 + (ml >> 4), px = x0 + (ml & 15);
        const bool v = py < p.H && px < p.W;
        ox[j] = v ? clamp_coord(__ldg(p.coords + (static_cast<size_t>(b) * 2 + 0) * HW + py * p.W + px)) : __int_as_float(0x7fc00000);
        oy[j] = v ? clamp_coord(__ldg(p.coords + (static_cast<size_t>(b) * 2 + 1) * HW + py * p.W + px)) : 0.f;
      }
    };
    // record of one unit: union box origin, rows needed, overflow
    struct Rec { int bx0, by0, nrows, ov; };
    auto make_rec = [&](int unit, const float (&ux)[4], const float (&uy)[4]) {
      const int l = unit_id(unit) / ntiles;
      const float inv = 1.f / static_cast<float>(1 << l);
      int lx0 = 0x7fffffff, ly0 = 0x7fffffff, lx1 = -0x7fffffff, ly1 = -0x7fffffff;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int ix0, iy0;
        if (ux[j] == ux[j] && window_origin(ux[j], uy[j], inv, p.H >> l, p.W >> l, ix0, iy0)) {   // NaN marks a pixel outside the image
          lx0 = min(lx0, ix0); ly0 = min(ly0, iy0); lx1 = max(lx1, ix0 + kG - 1); ly1 = max(ly1, iy0 + kG - 1);
        }
      }
      lx0 = __reduce_min_sync(0xffffffffu, lx0); ly0 = __reduce_min_sync(0xffffffffu, ly0);
      lx1 = __reduce_max_sync(0xffffffffu, lx1); ly1 = __reduce_max_sync(0xffffffffu, ly1);
      const bool any = lx1 >= lx0;
      Rec r;
      r.ov = any && (lx1 - lx0 + 1 > box_w(l) || ly1 - ly0 + 1 > box_h(l)) ? 1 : 0;
      r.bx0 = any ? lx0 : 0; r.by0 = any ? ly0 : 0;
#ifdef RNC_PROBE_NOTRIM
      r.nrows = any ? box_h(l) : 0;
#else
      r.nrows = any ? min((ly1 - ly0 + 2) & ~1, box_h(l)) : 0;
#endif
      return r;
    };
    int n = 0;
    auto publish = [&](int unit, const Rec& r) {
      const int slot = n & (kTiRing - 1);
      mbar_wait(&ti_empty[slot], ((n >> kTiShift) & 1) ^ 1);
      if (lane == 0) {
        ti[slot].bx0 = r.bx0; ti[slot].by0 = r.by0; ti[slot].overflow = r.ov; ti[slot].nrows = r.nrows;
        const int l = unit_id(unit) / ntiles;
        p.flags[(unit_id(unit) - l * ntiles) * kLevels + l] = r.ov;
        if (r.ov) *ov_count += 1;
        mbar_arrive(&ti_full[slot]);         // release: the unit record is visible to the waiters
      }
      __syncwarp();
      ++n;
    };
    // The record of unit k+1 is computed and published while unit k's loads are in flight (after its first kStages B
    // loads, when the ring is full and this warp would only wait), from coords fetched one unit earlier: nothing but
    // barrier waits sits between the last load of a unit and the first load of the next.
    // (a CTA's units are consecutive k: tile-major rounds first, then at most a run of remainder rounds)
    int unit = unit_at(0, ntiles, sched), next = -1;
    Rec cur = {0, 0, 0, 0}, nxt = {0, 0, 0, 0};
    int kfirst = 0;
    while (unit < 0 && kfirst + 1 < rounds) unit = unit_at(++kfirst, ntiles, sched);
    auto next_unit = [&](int& kk) {              // first valid unit after round kk (rounds with no unit for this CTA are skipped)
      int u = -1;
      while (u < 0 && kk + 1 < rounds) u = unit_at(++kk, ntiles, sched);
      return u;
    };
    int knext = kfirst;
    if (unit >= 0) {
      load_coords(unit_id(unit) % ntiles, pcx, pcy);
      cur = make_rec(unit, pcx, pcy);
      publish(unit, cur);
      next = next_unit(knext);
      if (next >= 0) load_coords(unit_id(next) % ntiles, pcx, pcy);
    }
    int it = 0, na = 0;
    while (unit >= 0) {
      const bool tm = unit_tm(unit);
      const int l = unit_id(unit) / ntiles, tile = unit_id(unit) - l * ntiles;
      const int b = tile / tpi, tr = tile - b * tpi;
      const int y0 = (tr / p.tiles_x) * kTY, x0 = (tr % p.tiles_x) * kTX;
      const int bx0 = cur.bx0, by0 = cur.by0, nrows = cur.nrows;
      const bool skip = cur.ov || nrows == 0;    // overflow: the exact kernel recomputes this tile; no rows: all-zero level
      const bool load_a = tm ? l == 0 : !skip;
      if (load_a) {
#pragma unroll
        for (int kb = 0; kb < kKB; ++kb) {
          mbar_wait(&a_empty[kb], (na & 1) ^ 1);
          if (elect_one()) {
#ifdef RNC_PROBE_NOLOAD
            mbar_arrive(&a_full[kb]);
#else
            mbar_expect_tx(&a_full[kb], kATile);
#if RNC_LOOKUP_L2HINT
            tma_load_4d_hint(sA + kb * kATile, &mF1, &a_full[kb], kb * 64, x0, y0, b, pol_keep);
#else
            tma_load_4d(sA + kb * kATile, &mF1, &a_full[kb], kb * 64, x0, y0, b);
#endif
#endif
          }
          __syncwarp();
        }
        ++na;
      }
      bool ahead_done = false;
      int after = -1;
      auto ahead = [&]() {
        if (ahead_done) return;
        ahead_done = true;
        if (next >= 0) {
          nxt = make_rec(next, pcx, pcy);
          publish(next, nxt);
          after = next_unit(knext);
          if (after >= 0) load_coords(unit_id(after) % ntiles, pcx, pcy);
        }
      };
      if (!skip) {
        // B: only the box rows some window needs, as 2-row TMA boxes (2 * bw positions = a multiple of 1024 bytes of the
        // SWIZZLE_128B stage, so the pieces tile the stage exactly as one big box would)
        const CUtensorMap* map = &mLv.m[l][0];
        const int cr = chunk_rows(l), bw = box_w(l);
        int issued = 0;
        for (int c = 0; c * cr < nrows; ++c) {
          const int rows = min(cr, nrows - c * cr);
          for (int kb = 0; kb < kKB; ++kb, ++it) {
            const int s = it % kStages, ph = (it / kStages) & 1;
            mbar_wait(&b_empty[s], ph ^ 1);
            if (elect_one()) {
#ifdef RNC_PROBE_NOLOAD
              mbar_arrive(&b_full[s]);
#else
              mbar_expect_tx(&b_full[s], rows * bw * 128);
#if RNC_LOOKUP_L2HINT
              tma_load_4d_hint(sB + s * kBStage, map + (rows >> 1) - 1, &b_full[s], kb * 64, bx0, by0 + c * cr, b, pol_keep);
#else
              tma_load_4d(sB + s * kBStage, map + (rows >> 1) - 1, &b_full[s], kb * 64, bx0, by0 + c * cr, b);   // rows is even
#endif
#endif
            }
            __syncwarp();
            if (++issued == kStages) ahead();
          }
        }
      }
      ahead();
      unit = next; cur = nxt; next = after;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loops, one elected lane issues)
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
    int it = 0, ch = 0, n = 0, na = 0;
    for (int k = 0; k < rounds; ++k) {
      const int unit = unit_at(k, ntiles, sched);
      if (unit < 0) continue;
      const bool tm = unit_tm(unit);
      const int l = unit_id(unit) / ntiles;
      const int slot = n & (kTiRing - 1);
      mbar_wait(&ti_full[slot], (n >> kTiShift) & 1);
      const int ov = ti[slot].overflow, nrows = ti[slot].nrows;
      __syncwarp();
      if (lane == 0) mbar_arrive(&ti_empty[slot]);
      ++n;
      const bool skip = ov || nrows == 0;
      const bool load_a = tm ? l == 0 : !skip, free_a = tm ? l == kLevels - 1 : !skip;
      if (skip) {                            // tile-major order: the A tile's barriers still turn over once per tile
        if (load_a) {
          for (int kb = 0; kb < kKB; ++kb) mbar_wait(&a_full[kb], na & 1);
          ++na;
        }
        if (free_a) {
          if (elect_one()) {
            for (int kb = 0; kb < kKB; ++kb) umma_commit(&a_empty[kb]);
          }
          __syncwarp();
        }
        continue;
      }
      const int cr = chunk_rows(l), nc = (nrows + cr - 1) / cr;
      for (int c = 0; c < nc; ++c, ++ch) {
        const bool first = c == 0, final = c == nc - 1;
        // instruction descriptor: D = F32, A = B = F16, K-major, N = positions of the rows this chunk holds, M = 128
        const uint32_t nn = min(cr, nrows - c * cr) * box_w(l);
        const uint32_t idesc = (1u << 4) | ((nn >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
        const int buf = ch & 1, use = ch >> 1;
        mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
        tcgen05_fence_after();
        for (int kb = 0; kb < kKB; ++kb, ++it) {
          const int s = it % kStages, ph = (it / kStages) & 1;
          if (first && load_a) mbar_wait(&a_full[kb], na & 1);
          mbar_wait(&b_full[s], ph);
          tcgen05_fence_after();
          const uint64_t ad = smem_desc_sw128(a_base + kb * kATile), bd = smem_desc_sw128(b_base + s * kBStage);
          if (elect_one()) {
#pragma unroll
#ifndef RNC_PROBE_NOMMA
            for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base + buf * 256, ad + 2 * kk, bd + 2 * kk, idesc, (kb | kk) != 0);
#endif
            umma_commit(&b_empty[s]);
            if (final && free_a) umma_commit(&a_empty[kb]);   // this K block of A is free for the next tile
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&acc_full[buf]);
        __syncwarp();
      }
      if (load_a) ++na;
    }
  } else {
    // ------------------------------------------------------------------ epilogue: gather + bilinear blend + store
    // thread = pixel = TMEM lane
    const int lg = warp & 3, ml = lg * 32 + lane;
    const int team = RNC_LOOKUP_EPI2 ? (warp - 2) >> 2 : 0;
    auto load_coord = [&](int tile, float& ox, float& oy, bool& v) {
      const int b = tile / tpi, tr = tile - b * tpi;
      const int py = (tr / p.tiles_x) * kTY + (ml >> 4), px = (tr % p.tiles_x) * kTX + (ml & 15);
      v = py < p.H && px < p.W;
      ox = v ? clamp_coord(__ldg(p.coords + (static_cast<size_t>(b) * 2 + 0) * HW + py * p.W + px)) : 0.f;
      oy = v ? clamp_coord(__ldg(p.coords + (static_cast<size_t>(b) * 2 + 1) * HW + py * p.W + px)) : 0.f;
    };
    float cx = 0.f, cy = 0.f, ncx = 0.f, ncy = 0.f;
    bool valid = false, nvalid = false;
#ifndef RNC_PROBE_NOEPI
    if (team == 0) {                           // the 7 pad channels of this pixel's output rows stay zero for the whole kernel
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(stage + ml * kLvlStride + 80) = z;
      *reinterpret_cast<uint4*>(stage + (128 + ml) * kLvlStride + 80) = z;
    }
#endif
    int unit = unit_at(0, ntiles, sched);
    if (unit >= 0) load_coord(unit_id(unit) % ntiles, ncx, ncy, nvalid);
    int ch = 0, n = 0;
    for (int k = 0; k < rounds; ++k) {
      const int cur = unit;
      cx = ncx; cy = ncy; valid = nvalid;
      unit = k + 1 < rounds ? unit_at(k + 1, ntiles, sched) : -1;
      if (unit >= 0) load_coord(unit_id(unit) % ntiles, ncx, ncy, nvalid);
      if (cur < 0) continue;
      const int l = unit_id(cur) / ntiles, tile = unit_id(cur) - l * ntiles;
      const int b = tile / tpi, tr = tile - b * tpi;
      const int y0 = (tr / p.tiles_x) * kTY, x0 = (tr % p.tiles_x) * kTX;
      const int slot = n & (kTiRing - 1);
      mbar_wait(&ti_full[slot], (n >> kTiShift) & 1);
      if (!ti[slot].overflow) {
        EpiCtx c{p, &ti[slot], scratch, stage, acc_full, acc_empty, tmem_base, b, y0, x0, lg, ml, lane, valid, cx, cy, team};
        lookup_epilogue(c, ch, l, &mOutHi, &mOutLo);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ti_empty[slot]);
      ++n;
    }
    if (lane == 0 && team == 0) bulk_wait_all0();   // the output tile must outlive its TMA stores
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  // ---- units of this CTA that overflowed their box (flagged by this CTA's own producer warp): exact path, all warps
  if (*ov_count == 0) return;                  // the common case: nothing to redo
  for (int k = 0; k < rounds; ++k) {
    const int unit = unit_at(k, ntiles, sched);
    if (unit < 0) continue;
    const int l = unit_id(unit) / ntiles, tile = unit_id(unit) - l * ntiles;
    if (p.flags[tile * kLevels + l] != 0) exact_unit(p, l, tile, scratch, warp, lane, kThreads / 32);
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
  return static_cast<size_t>(B) * ((H + kTY - 1) / kTY) * ((W + kTX - 1) / kTX) * kLevels * sizeof(int);
}

static int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
  }
  return cached;
}

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
  if (!aligned16(f1h_cl) || !aligned16(f2h_pyr) || !aligned16(workspace) || !aligned16(out_hi) || !aligned16(out_lo)) return RNC_ERR_BAD_POINTER;
  if (workspace_bytes < rnc_corr_lookup_umma_workspace_bytes(B, H, W)) return RNC_ERR_WORKSPACE;
  if (!encode_fn()) return RNC_ERR_UNSUPPORTED;

  Params p;
  p.f1_cl = f1_cl; p.f2_pyr = f2_pyr;
  p.coords = coords;
  p.out_hi = static_cast<__half*>(out_hi); p.out_lo = static_cast<__half*>(out_lo); p.ldo = ldo;
  p.flags = static_cast<int*>(workspace);
  p.B = B; p.H = H; p.W = W;
  p.tiles_x = (W + kTX - 1) / kTX; p.tiles_y = (H + kTY - 1) / kTY;
  p.scale = 1.0f / sqrtf(static_cast<float>(D));
  {
    // Hybrid order by default (see unit_at); RNC_LOOKUP_SCHED=level|tile select the pure orders (developer override).
    static const char* env = getenv("RNC_LOOKUP_SCHED");
    p.tile_major = env == nullptr ? 2 : env[0] == 't' ? 1 : env[0] == 'l' ? 0 : 2;
  }

  CUtensorMap maps[7];
  LevelMaps lv;
  bool ok = make_act_map(&maps[0], f1h_cl, kD, kD, B, H, W, kTX, kTY);
  // output planes [B][H][W][ldo] halves: un-swizzled store boxes of one level (88 channels) x 16 x 2 pixels
  ok = ok && make_plain_map(&maps[5], out_hi, ldo, ldo, B, H, W, kLvlStride, kTX, 2);
  ok = ok && make_plain_map(&maps[6], out_lo, ldo, ldo, B, H, W, kLvlStride, kTX, 2);
  size_t off = 0;
  for (int l = 0; l < kLevels; ++l) {
    const int Hl = H >> l, Wl = W >> l;
    for (int q = 0; q < kBoxKinds; ++q)
      ok = ok && make_act_map(&lv.m[l][q], static_cast<const __half*>(f2h_pyr) + off, kD, kD, B, Hl, Wl, box_w(l), 2 * (q + 1));
    off += static_cast<size_t>(B) * Hl * Wl * kD;
  }
  if (!ok) return RNC_ERR_BAD_SHAPE;

  static unsigned long long done = 0;
  if (int st = ensure_dyn_smem(corr_lookup_umma_kernel, kSmemTotal, &done)) return st;
  const int ntiles = B * p.tiles_x * p.tiles_y;
  const int grid = ntiles * kLevels < sm_count() ? ntiles * kLevels : sm_count();   // persistent: one CTA per SM
  cudaError_t e = launch_pdl(corr_lookup_umma_kernel, dim3(grid), dim3(kThreads), kSmemTotal, as_stream(stream), maps[0], lv, maps[5], maps[6], p);
  if (e != cudaSuccess) { g_last_cuda_error = static_cast<int>(e); return RNC_ERR_CUDA; }
  return after_launch();
}
