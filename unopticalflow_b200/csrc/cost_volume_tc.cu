// a1 backward on the 5th-generation tensor cores: cost-volume gradient as a banded 3xTF32 GEMM (tcgen05 + TMEM).
// OPT-IN (UOF_CV_TC=1): parity-green but NOT faster than the CUDA-core kernel (0.6-0.8x at C = 32, 1.0-1.2x at C = 64/96) --
// measurements, pipeline trace and the reasons are in profiles/r2_tc_cost_volume_bwd.md; the default path stays
// cost_volume_tma.cu.  Kept because it is the evidence behind "0.6 of the HBM roofline is unreachable for a1-bwd".
// Reference math: PWC_tf.corr_naive, /root/reference/core/networks/structures/pwc_tf.py:97-106
//   out[b, 9i+j, y, x] = (1/C) sum_c f1[b,c,y,x] * f2[b,c,y+i-4,x+j-4]
//   gf1[c,p] = (1/C) sum_d gout[d,p]        * f2[c,p+disp(d)]            ("role 0")
//   gf2[c,q] = (1/C) sum_d gout[d,q-disp(d)] * f1[c,q-disp(d)]            ("role 1"),   disp(9i+j) = (i-4, j-4)
//
// Formulation.  A CTA owns a 16x8 pixel tile (M = 128 = the TMEM lanes) and all C channels (N = C).  The neighbourhood of
// the tile is a 24x16 window (K = 384, index k = qy*16 + qx).  For both roles
//   D_r[pixel m, channel c] = sum_k A_r[m, k] * F_r[c, window pixel k]
// where F_0 = f2, F_1 = f1 over the SAME window and A_r is the band matrix of the gradient:
//   lane m = (ty,tx), column k = (qy,qx) is non-zero iff 0 <= qy-ty <= 8 and 0 <= qx-tx <= 8; with d1 = 9(qy-ty) + (qx-tx)
//   A_0[m,k] = gout[d1, tile pixel m]                                    (g1: the tile of gout)
//   A_1[m,k] = gout[80-d1, tile pixel m - disp(80-d1)] =: g2[80-d1][m]   (g2: the "sheared" tile, gathered with zero-filling
//              4-byte cp.async: the TMA unit rejects box origins that are not 16-byte aligned, tools/tc_probe.cu)
// A_r lives in TENSOR MEMORY: converter warps build it one window row (16 columns) at a time from shared memory
// (bank-conflict free: the address is (const - lane terms)*128 + lane) and write it with tcgen05.st, so the band matrix
// never touches shared memory.  B = F_r is staged in shared memory in the K-major SWIZZLE_64B canonical UMMA layout
// ([channel][16 floats], 64-byte rows).  fp32 parity (1e-4) needs 3xTF32: the tensor core TRUNCATES fp32 inputs to tf32
// (measured), so hi = x & 0xFFFFE000 is what it sees of the raw tile, lo = x - hi is computed by the converter warps (A: in
// registers on the way to TMEM; B: one pass over the tile), and D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo.
// One elected thread issues the tcgen05.mma's (M = 128, K = 8); mbarrier rings couple loader -> converters -> MMA issuer ->
// (tcgen05.commit) -> back.  Epilogue: tcgen05.ld of the accumulators, scale by 1/C, 32-byte-segment stores to gf1 / gf2.
#include <stdlib.h>

#include "cost_volume.h"
#include "tcgen05.cuh"

namespace uof {
namespace cv {
namespace {

constexpr int TH = 16, TW = 8;                  // pixel tile (M = 128)
constexpr int QH = TH + 2 * RAD, QW = TW + 2 * RAD;   // 24 x 16 window
constexpr int NDISP = ND * ND;                  // 81
constexpr int kGroups = 3;                      // converter groups (4 warps each), window row c is built by group c % kGroups
constexpr int kConvThreads = 128 * kGroups;
constexpr int kTcThreads = 64 + kConvThreads;   // warp 0: TMA producer, warp 1: MMA issuer, warps 2..: converters / epilogue
constexpr int kGBytes = NDISP * TH * TW * 4;    // one gout tile: 41 472 B

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int C>
struct TcCfg {
  // A "stage" is one window row (16 columns of K) of BOTH roles: two band-matrix chunks in TMEM ({hi, lo} x 16 columns each)
  // and two feature tiles in shared memory (hi by TMA, lo by the converters).
  //
  // What bounds the kernel (ncu, profiles/r2_tc_bwd_*):
  //  * A tcgen05.mma with A in TMEM re-reads its 128 x 8 x 4 B A tile for every instruction while the math of an M=128,
  //    N=32, K=8 tf32 MMA is only 16 cycles.  For C <= 64 the hi and lo halves of a feature tile therefore sit in ONE
  //    shared-memory operand of 2C rows ([B_hi ; B_lo]): A_hi * [B_hi ; B_lo]^T is one MMA of N = 2C into accumulator columns
  //    [0, 2C), A_lo * B_hi^T a second one of N = C into [0, C), and the epilogue adds the halves: two A reads per K-step
  //    instead of three.
  //  * The converters.  With four converter warps (one per scheduler, nothing to switch to) every LDS / ALU dependency of
  //    the ~300-instruction chunk construction was exposed and a row took ~1000 cycles whatever the ring depth (v1-v4:
  //    0.64-0.78x of the CUDA-core kernel with the tensor pipe < 40 % busy).  Three converter groups work on three
  //    different rows at a time (three warps per scheduler) and both roles of a row share the band predicates and addresses.
  static constexpr bool kStack = C <= 64;
  static constexpr int kDCols = kStack ? 2 * C : C;                    // accumulator columns per role
  static constexpr int kACols = 64;                                    // per stage: role x {hi, lo} x 16 columns
  //  * TMA latency.  A feature row is 2 x C box rows of 64 bytes, one per channel plane: ~2-3 us from issue to the
  //    mbarrier under load (v5: the converters spent 20 % of their time waiting for it with a 6-deep ring).  The shared-memory
  //    ring (kNB rows of {[B_hi ; B_lo]} x 2 roles, freed by the MMA commit) is therefore deeper than the TMEM ring of band
  //    chunks (kNS), which only bridges converters -> MMA.
  static constexpr int kNS = (512 - 2 * kDCols) / kACols < 6 ? (512 - 2 * kDCols) / kACols : 6;      // TMEM ring depth
  static constexpr int kNB = C <= 32 ? 12 : (C <= 64 ? 8 : (C <= 96 ? 5 : 4));                       // smem ring depth
  static constexpr int kTileBytes = C * QW * 4;                        // one [C][16] operand tile (C * 64 B, multiple of 512)
  static constexpr int kStageBytes = 4 * kTileBytes;                   // role x [hi C rows ; lo C rows]
  static constexpr int kTmemCols = 512;
  static constexpr size_t kSmem = 1024 + 2 * (size_t)kGBytes + (size_t)kNB * kStageBytes + 1024;
  static_assert(C % 16 == 0 && C >= 16 && C <= 128, "N = C must be a multiple of 16 (M = 128) and fit TMEM");
  static_assert(kNS >= 3 && kNB >= kNS && 2 * kDCols + kNS * kACols <= 512, "TMEM budget");
  static_assert(kSmem <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ void conv_bar_sync() {      // named barrier over the converter warps only
  asm volatile("bar.sync 1, %0;" ::"n"(kConvThreads) : "memory");
}

template <int C>
__global__ void __launch_bounds__(kTcThreads, 1)
cost_volume_bwd_tc_kernel(const float* __restrict__ gout, long long gout_bs, const float* __restrict__ f1,
                          const float* __restrict__ f2, float* __restrict__ gf1, float* __restrict__ gf2, int H, int W,
                          float inv_c, long long* __restrict__ trace) {
  // trace != nullptr (UOF_CV_TC_TRACE, tools/cv_tc_trace.py): CTA (0,0,0) records clock64() at every pipeline event of every
  // window row -- [row][event]: 0 loader issued, 1 loader published, 2 converter got its TMEM stage, 3 band chunks stored,
  // 4 feature row seen, 5 ready signalled, 6 MMA issuer got the row, 7 MMAs issued + committed; rows 24.. hold
  // {prologue done, accumulators complete, epilogue done}.
  const bool tr = trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define UOF_TRACE(row, ev)                                   \
  do {                                                       \
    if (tr) trace[(row) * 8 + (ev)] = clock64();             \
  } while (0)
  using Cfg = TcCfg<C>;
  constexpr int NS = Cfg::kNS, NB = Cfg::kNB;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte aligned carve-up (the operand tiles need 512 for SWIZZLE_64B)
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* g1 = reinterpret_cast<float*>(base);                          // [81][128]  gout tile
  float* g2 = g1 + NDISP * TH * TW;                                    // [81][128]  sheared gout tile
  unsigned char* ring = base + 2 * kGBytes;                            // [NB][role]{[C][16] raw feature row (TMA) ; [C][16] lo}
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + NB * Cfg::kStageBytes);
  uint64_t* bar_bfull = bars + 1;              // [NB] feature rows landed
  uint64_t* bar_bfree = bar_bfull + NB;        // [NB] MMAs that read the smem stage retired
  uint64_t* bar_ready = bar_bfree + NB;        // [NS] converters done (A chunks in TMEM, B_lo in smem)
  uint64_t* bar_free = bar_ready + NS;         // [NS] MMAs that read the TMEM stage retired
  uint64_t* bar_done = bar_free + NS;          // all MMAs retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NB; ++s) {
      mbar_init(bar_bfull + s, 1);
      mbar_init(bar_bfree + s, 1);
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(bar_ready + s, 4);
      mbar_init(bar_free + s, 1);
    }
    mbar_init(bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: [0, kDCols) role 0 accumulators, [kDCols, 2 kDCols) role 1, then NS stages {A0 hi, A0 lo, A1 hi, A1 lo} x 16
  const uint32_t tmem_a = tmem + 2 * Cfg::kDCols;

  if (warp == 0) {
    // ============================================ feature-row loader (one warp) ============================================
    // 16-byte zero-filling cp.async into the K-major SWIZZLE_64B canonical layout: row = channel (64 bytes = the 16 window
    // columns), the 16-byte chunk index XORed with (channel >> 1) & 3.  W % 4 == 0 and x0 % 8 == 0, so a chunk is entirely
    // inside or entirely outside the image.  NOT the TMA unit: a {16 x, C channels, 1 row} box is C box rows of 64 bytes one
    // channel plane apart, and the TMA engine needed ~25 cycles per such row -- 1536 rows = 38k cycles per tile, which made
    // v1-v6 of this kernel slower than the CUDA-core one whatever else was tuned (profiles/r2_tc_bwd_*).
    constexpr int kLag = 3;                       // rows in flight per lane before the oldest is published
    static_assert(NB > kLag, "ring too shallow for the loader lag");
    constexpr int kChunks = C * 4;                // 16-byte chunks per [C][16] tile
    const size_t plane = (size_t)H * W;
    const float* src1 = f1 + (size_t)b * C * plane;
    const float* src2 = f2 + (size_t)b * C * plane;
    for (int c = 0; c < QH + kLag; ++c) {
      if (c < QH) {
        const int sb = c % NB;
        if (c >= NB) mbar_wait(bar_bfree + sb, ((c / NB) - 1) & 1);
        const int yy = y0 - RAD + c;
        const bool rowok = yy >= 0 && yy < H;
        unsigned char* st = ring + sb * Cfg::kStageBytes;
#pragma unroll
        for (int i = 0; i < kChunks / 32; ++i) {
          const int k = lane + 32 * i, ch = k >> 2, part = k & 3;
          const int xx = x0 - RAD + 4 * part;
          const bool ok = rowok && xx >= 0 && xx < W;
          const size_t off = (size_t)ch * plane + (size_t)(rowok ? yy : 0) * W + (ok ? xx : 0);
          float* dst = reinterpret_cast<float*>(st + ch * 64 + ((part ^ ((ch >> 1) & 3)) << 4));
          // role 0 (gf1) contracts with f2, role 1 (gf2) with f1; same window row for both
          cp_async_16(dst, src2 + off, ok);
          cp_async_16(dst + 2 * Cfg::kTileBytes / 4, src1 + off, ok);
        }
      }
      cp_async_commit();
      if (lane == 0 && c < QH) UOF_TRACE(c, 0);
      if (c >= kLag) {
        cp_async_wait<kLag>();
        tc::fence_proxy_async();                  // the tensor core reads these tiles through the async proxy
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_bfull + (c - kLag) % NB);
          UOF_TRACE(c - kLag, 1);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================== MMA issuer ==================================================
    if (lane == 0) {
      constexpr uint32_t idesc_c = tc::idesc_tf32(128, C), idesc_2c = tc::idesc_tf32(128, Cfg::kStack ? 2 * C : C);
      for (int c = 0; c < QH; ++c) {
        const int s = c % NS, sb = c % NB;
        mbar_wait(bar_bfull + sb, (c / NB) & 1);
        mbar_wait(bar_ready + s, (c / NS) & 1);
        tc::fence_after_sync();
        UOF_TRACE(c, 6);
        const uint32_t st = smem_u32(ring + sb * Cfg::kStageBytes);
        const uint32_t ta = tmem_a + s * Cfg::kACols;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t hi = st + 2 * r * Cfg::kTileBytes, lo = hi + Cfg::kTileBytes;
          const uint32_t d = tmem + r * Cfg::kDCols;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t bh = tc::smem_desc_kmajor(hi + 32 * ks, 512, tc::kLayoutSw64);
            const uint32_t ah = ta + 32 * r + 8 * ks, al = ah + 16;
            const bool acc = c > 0 || ks > 0;
            if (Cfg::kStack) {
              tc::mma_tf32_ts(d, ah, bh, idesc_2c, acc);             // A_hi * [B_hi ; B_lo]  -> columns [0, 2C)
              tc::mma_tf32_ts(d, al, bh, idesc_c, true);             // A_lo * B_hi           -> columns [0, C)
            } else {
              const uint64_t bl = tc::smem_desc_kmajor(lo + 32 * ks, 512, tc::kLayoutSw64);
              tc::mma_tf32_ts(d, ah, bh, idesc_c, acc);              // hi * hi
              tc::mma_tf32_ts(d, al, bh, idesc_c, true);             // lo * hi
              tc::mma_tf32_ts(d, ah, bl, idesc_c, true);             // hi * lo
            }
          }
        }
        tc::mma_commit(bar_free + s);      // both stages (TMEM chunks, smem tiles) are released when these MMAs retire
        tc::mma_commit(bar_bfree + sb);
        UOF_TRACE(c, 7);
      }
      tc::mma_commit(bar_done);
    }
  } else {
    // ========================================= converters (3 x 4 warps) + epilogue =========================================
    const int cw = warp - 2;                       // converter warp 0..11
    const int grp = cw >> 2;                       // group 0..2
    const int q = warp & 3;                        // TMEM lane quadrant this warp may touch
    const int m = q * 32 + lane;                   // tile pixel = TMEM lane
    const int ty = m >> 3, tx = m & 7;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int gt = (cw & 3) * 32 + lane;           // 0..127: thread index inside the group (B pass)
    {
      // sheared gout tile g2[d][m] = gout[b, d, y - (i-4), x - (j-4)] (zero outside the image), displacement rows split over the
      // groups: zero-filling 4-byte cp.async (the TMA unit rejects box origins that are not 16-byte aligned)
      const int y = y0 + ty, x = x0 + tx;
      const float* gb = gout + (size_t)b * gout_bs;
      const size_t plane = (size_t)H * W;
#pragma unroll
      for (int ii = 0; ii < ND / kGroups; ++ii) {
        const int i = ii * kGroups + grp;
        const int yy = y - (i - RAD);
        const bool yok = yy >= 0 && yy < H;
#pragma unroll
        for (int j = 0; j < ND; ++j) {
          const int xx = x - (j - RAD);
          const bool ok = yok && xx >= 0 && xx < W;
          const int d = i * ND + j;
          cp_async_4(g2 + d * (TH * TW) + m, ok ? gb + d * plane + (size_t)yy * W + xx : gb, ok);
        }
      }
      // straight gout tile g1[d][m] = gout[b, d, y0 + ty, x0 + tx]: 81 x 16 rows of two 16-byte chunks (x0 % 8 == 0, W % 4 == 0)
      const int ctid = cw * 32 + lane;
      for (int k = ctid; k < NDISP * TH * 2; k += kConvThreads) {
        const int d = k >> 5, rem = k & 31, row = rem >> 1, half = rem & 1;
        const int yy = y0 + row, xx = x0 + 4 * half;
        const bool ok = yy < H && xx < W;
        cp_async_16(g1 + d * (TH * TW) + row * TW + 4 * half, ok ? gb + d * plane + (size_t)yy * W + xx : gb, ok);
      }
      cp_async_commit();
      cp_async_wait<0>();
    }
    conv_bar_sync();                               // every group reads all of g1 / g2
    const bool trc = (cw & 3) == 0 && lane == 0;   // first warp of each group records
    if (cw == 0 && lane == 0) UOF_TRACE(QH, 0);
    const uint32_t xmask = 0x1FFu << tx;           // window columns qx with 0 <= qx - tx <= 8
    for (int c = grp; c < QH; c += kGroups) {
      const int s = c % NS, sb = c % NB;
      if (c >= NS) {
        mbar_wait(bar_free + s, ((c / NS) - 1) & 1);
        tc::fence_after_sync();
      }
      if (trc) UOF_TRACE(c, 2);
      // ---- band chunks of window row qy = c, both roles -> TMEM
      const uint32_t ta = tmem_a + s * Cfg::kACols + lane_addr;
      const int dyi = c - ty;
      const uint32_t msk = (dyi >= 0 && dyi <= 2 * RAD) ? xmask : 0u;
      if (__any_sync(kFullMask, msk != 0u)) {
        const int d_base = dyi * ND - tx;          // d1 = d_base + qx
        const float* p0 = g1 + d_base * (TH * TW) + m;                       // role 0: g1[d1][m]
        const float* p1 = g2 + (NDISP - 1 - d_base) * (TH * TW) + m;         // role 1: g2[80 - d1][m]
        uint32_t h[16], l[16];
#pragma unroll
        for (int qx = 0; qx < QW; ++qx) {
          const float v = ((msk >> qx) & 1u) ? p0[qx * (TH * TW)] : 0.0f;
          h[qx] = tc::tf32_hi(v);
          l[qx] = tc::tf32_lo(v, h[qx]);
        }
        tc::tmem_st16(ta, h);
        tc::tmem_st16(ta + 16, l);
#pragma unroll
        for (int qx = 0; qx < QW; ++qx) {
          const float v = ((msk >> qx) & 1u) ? p1[-qx * (TH * TW)] : 0.0f;
          h[qx] = tc::tf32_hi(v);
          l[qx] = tc::tf32_lo(v, h[qx]);
        }
        tc::tmem_st16(ta + 32, h);
        tc::tmem_st16(ta + 48, l);
      } else {
        uint32_t z[16];
#pragma unroll
        for (int qx = 0; qx < QW; ++qx) z[qx] = 0u;
        tc::tmem_st16(ta, z);
        tc::tmem_st16(ta + 16, z);
        tc::tmem_st16(ta + 32, z);
        tc::tmem_st16(ta + 48, z);
      }
      // ---- B_lo = B - trunc_tf32(B) for both feature rows (same swizzled position in the lo half)
      if (trc) UOF_TRACE(c, 3);
      mbar_wait(bar_bfull + sb, (c / NB) & 1);
      if (trc) UOF_TRACE(c, 4);
      {
        float4* st = reinterpret_cast<float4*>(ring + sb * Cfg::kStageBytes);
        constexpr int kVec = Cfg::kTileBytes / 16;                     // float4 per tile = 4 C
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
          for (int i = gt; i < kVec; i += 128) {
            const float4 v = st[2 * r * kVec + i];
            float4 o;
            o.x = __uint_as_float(tc::tf32_lo(v.x, tc::tf32_hi(v.x)));
            o.y = __uint_as_float(tc::tf32_lo(v.y, tc::tf32_hi(v.y)));
            o.z = __uint_as_float(tc::tf32_lo(v.z, tc::tf32_hi(v.z)));
            o.w = __uint_as_float(tc::tf32_lo(v.w, tc::tf32_hi(v.w)));
            st[(2 * r + 1) * kVec + i] = o;
          }
        }
      }
      tc::fence_proxy_async();
      tc::wait_st();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ready + s);
      if (trc) UOF_TRACE(c, 5);
    }
    // ---- epilogue: D_r[pixel][channel] -> g{f1,f2}[b][channel][y][x] * (1/C); 16-channel chunks dealt round robin to the groups
    mbar_wait(bar_done, 0);
    tc::fence_after_sync();
    if (cw == 0 && lane == 0) UOF_TRACE(QH, 1);
    const int y = y0 + ty, x = x0 + tx;
    const bool live = y < H && x < W;
    const size_t plane = (size_t)H * W;
    constexpr int kChunks = 2 * (C / 16);
#pragma unroll
    for (int e = 0; e < (kChunks + kGroups - 1) / kGroups; ++e) {
      const int id = e * kGroups + grp;
      if (id < kChunks) {
        const int r = id / (C / 16), j = id - r * (C / 16);
        float* dst = (r == 0 ? gf1 : gf2) + (size_t)b * C * plane + (size_t)y * W + x;
        uint32_t v[16], w[16];
        tc::tmem_ld16(tmem + lane_addr + r * Cfg::kDCols + 16 * j, v);
        if (Cfg::kStack) tc::tmem_ld16(tmem + lane_addr + r * Cfg::kDCols + C + 16 * j, w);      // the A_hi * B_lo half
        tc::wait_ld();
        if (live) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float acc = __uint_as_float(v[k]);
            if (Cfg::kStack) acc += __uint_as_float(w[k]);
            dst[(size_t)(16 * j + k) * plane] = acc * inv_c;
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) UOF_TRACE(QH, 2);
#undef UOF_TRACE
  if (warp == 1) tc::tmem_dealloc(tmem, Cfg::kTmemCols);
}

long long* g_trace = nullptr;      // device buffer for UOF_CV_TC_TRACE (debug only)

template <int C>
int launch_bwd_tc(const float* gout, long long gout_bs, const float* f1, const float* f2, float* gf1, float* gf2, int B, int H,
                  int W, cudaStream_t stream) {
  auto kern = cost_volume_bwd_tc_kernel<C>;
  UOF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<C>::kSmem));
  kern<<<dim3(ceil_div(W, TW), ceil_div(H, TH), B), kTcThreads, TcCfg<C>::kSmem, stream>>>(gout, gout_bs, f1, f2, gf1, gf2, H, W,
                                                                                            1.0f / (float)C, g_trace);
  count_launch();
  return check_launch("cost_volume_bwd (tcgen05)");
}

}  // namespace

// Debug: allocate (once) and return the device trace buffer of (QH + 1) x 8 clock64 stamps; see the kernel.
extern "C" long long* uof_cv_tc_trace_buffer() {
  if (!g_trace) {
    if (cudaMalloc(&g_trace, (QH + 1) * 8 * sizeof(long long)) != cudaSuccess) return nullptr;
    cudaMemset(g_trace, 0, (QH + 1) * 8 * sizeof(long long));
  }
  return g_trace;
}

extern "C" int uof_cv_tc_trace_read(long long* host) {
  if (!g_trace) return 1;
  return cudaMemcpy(host, g_trace, (QH + 1) * 8 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 2;
}

// Tensor-core backward, opt-in: UOF_CV_TC=1 uses it where a level fills the machine, UOF_CV_FORCE_TC=1 wherever the shape
// allows.  Returns false when it is not selected or does not apply (C not in {32,64,96,128}, W % 4 != 0, unaligned pointers);
// otherwise launches and stores the status.
bool bwd_tc(const float* gout, long long gout_bs, const float* f1, const float* f2, float* gf1, float* gf2, int B, int C,
            int H, int W, cudaStream_t stream, int* rc) {
  const bool force = getenv("UOF_CV_FORCE_TC") != nullptr;     // read per call: tests and benches switch paths at run time
  const bool on = force || getenv("UOF_CV_TC") != nullptr;
  if (!on || !(C == 32 || C == 64 || C == 96 || C == 128)) return false;
  if (W % 4 != 0 || gout_bs % 4 != 0 || !aligned16(gout) || !aligned16(f1) || !aligned16(f2)) return false;
  if (B > 65535 || ceil_div(H, TH) > 65535) return false;
  const long long tiles = (long long)ceil_div(W, TW) * ceil_div(H, TH) * B;
  if (!force && tiles < 2 * kNumSMs) return false;          // the latency-bound small levels keep their own kernels
  switch (C) {
    case 32: *rc = launch_bwd_tc<32>(gout, gout_bs, f1, f2, gf1, gf2, B, H, W, stream); break;
    case 64: *rc = launch_bwd_tc<64>(gout, gout_bs, f1, f2, gf1, gf2, B, H, W, stream); break;
    case 96: *rc = launch_bwd_tc<96>(gout, gout_bs, f1, f2, gf1, gf2, B, H, W, stream); break;
    default: *rc = launch_bwd_tc<128>(gout, gout_bs, f1, f2, gf1, gf2, B, H, W, stream); break;
  }
  return true;
}

}  // namespace cv
}  // namespace uof
