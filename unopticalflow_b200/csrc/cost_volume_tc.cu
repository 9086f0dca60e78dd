// a1 backward on the 5th-generation tensor cores: cost-volume gradient as a banded 3xTF32 GEMM (tcgen05 + TMEM + TMA).
// Reference math: PWC_tf.corr_naive, /root/reference/core/networks/structures/pwc_tf.py:97-106
//   out[b, 9i+j, y, x] = (1/C) sum_c f1[b,c,y,x] * f2[b,c,y+i-4,x+j-4]
//   gf1[c,p] = (1/C) sum_d gout[d,p]        * f2[c,p+disp(d)]            ("role 0")
//   gf2[c,q] = (1/C) sum_d gout[d,q-disp(d)] * f1[c,q-disp(d)]            ("role 1"),   disp(9i+j) = (i-4, j-4)
//
// Why tensor cores: on CUDA cores this contraction is shared-memory-bandwidth bound at 0.26 of the HBM roofline / 0.30 of the
// FP32 peak and needs 70-87 % of that peak to reach 0.6.
//
// Formulation.  A CTA owns a 16x8 pixel tile (M = 128 = the TMEM lanes) and all C channels (N = C).  The neighbourhood of
// the tile is a 24x16 window (K = 384, index k = qy*16 + qx).  For both roles
//   D_r[pixel m, channel c] = sum_k A_r[m, k] * F_r[c, window pixel k]
// where F_0 = f2, F_1 = f1 over the SAME window and A_r is the band matrix of the gradient:
//   lane m = (ty,tx), column k = (qy,qx) is non-zero iff 0 <= qy-ty <= 8 and 0 <= qx-tx <= 8; with d1 = 9(qy-ty) + (qx-tx)
//   A_0[m,k] = gout[d1, tile pixel m]                                    (tile of gout, one TMA box load)
//   A_1[m,k] = gout[80-d1, tile pixel m - disp(80-d1)] =: g2[80-d1][m]   (the "sheared" tile: every lane copies the 81
//              values of ITS OWN column with zero-filling 4-byte cp.async -- the TMA unit rejects box origins that are not
//              16-byte aligned (measured, tools/tc_probe.cu: illegal instruction), so a per-plane shifted box is not an option)
// A_r lives in TENSOR MEMORY: the four converter warps build it one window row (16 columns) at a time from shared memory
// (bank-conflict free: the address is (const - lane terms)*128 + lane) and write it with tcgen05.st, so the band matrix
// never touches shared memory (in shared memory its construction + the tensor core's re-read would cost 7.5k cycles per
// tile at 128 B/clk -- more than the CUDA-core kernel).  B = F_r arrives by TMA straight from the NCHW tensors through a
// tensor map with permuted dimensions (x, c, y, b) and CU_TENSOR_MAP_SWIZZLE_64B: the box {16 x, C channels, 1 row} lands
// as the K-major canonical UMMA layout ([channel][16 floats], 64-byte rows), out-of-image elements zero-filled.
// fp32 parity (1e-4) needs 3xTF32: the tensor core TRUNCATES fp32 inputs to tf32 (measured, tools/tc_probe.cu), so
// hi = x & 0xFFFFE000 is what it sees of the raw tile, lo = x - hi is computed by the converter warps (A: in registers on the
// way to TMEM; B: one pass over the tile into a second buffer), and D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo.
// 24 window rows x 2 roles x 2 K-steps x 3 = 288 tcgen05.mma (M=128, N=C, K=8) per tile, accumulators in TMEM, one elected
// thread issues; a 3-4 stage mbarrier ring couples TMA producer -> converters -> MMA issuer -> (tcgen05.commit) -> back.
// Epilogue: tcgen05.ld of D_0 / D_1, scale by 1/C, 32-byte-segment coalesced stores to gf1 / gf2.
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
cost_volume_bwd_tc_kernel(const __grid_constant__ CUtensorMap gmap_tile, const __grid_constant__ CUtensorMap f1map,
                          const __grid_constant__ CUtensorMap f2map, const float* __restrict__ gout, long long gout_bs,
                          float* __restrict__ gf1, float* __restrict__ gf2, int H, int W, float inv_c) {
  using Cfg = TcCfg<C>;
  constexpr int NS = Cfg::kNS, NB = Cfg::kNB;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte aligned carve-up (the operand tiles need 512 for SWIZZLE_64B)
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* g1 = reinterpret_cast<float*>(base);                          // [81][128]  gout tile
  float* g2 = g1 + NDISP * TH * TW;                                    // [81][128]  sheared gout tile
  unsigned char* ring = base + 2 * kGBytes;                            // [NB][role]{[C][16] raw feature row (TMA) ; [C][16] lo}
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + NB * Cfg::kStageBytes);
  uint64_t* bar_g = bars;                      // gout tile landed
  uint64_t* bar_bfull = bars + 1;              // [NB] feature rows landed
  uint64_t* bar_bfree = bar_bfull + NB;        // [NB] MMAs that read the smem stage retired
  uint64_t* bar_ready = bar_bfree + NB;        // [NS] converters done (A chunks in TMEM, B_lo in smem)
  uint64_t* bar_free = bar_ready + NS;         // [NS] MMAs that read the TMEM stage retired
  uint64_t* bar_done = bar_free + NS;          // all MMAs retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&gmap_tile);
    tma_prefetch_desc(&f1map);
    tma_prefetch_desc(&f2map);
    mbar_init(bar_g, 1);
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
    // ================================================= TMA producer =================================================
    if (lane == 0) {
      mbar_expect_tx(bar_g, kGBytes);
      tma_load_4d(g1, &gmap_tile, bar_g, x0, y0, 0, b);
      for (int c = 0; c < QH; ++c) {
        const int sb = c % NB;
        if (c >= NB) mbar_wait(bar_bfree + sb, ((c / NB) - 1) & 1);
        unsigned char* st = ring + sb * Cfg::kStageBytes;
        mbar_expect_tx(bar_bfull + sb, 2 * Cfg::kTileBytes);
        // role 0 (gf1) contracts with f2, role 1 (gf2) with f1; same window row for both
        tma_load_4d(st, &f2map, bar_bfull + sb, x0 - RAD, 0, y0 - RAD + c, b);
        tma_load_4d(st + 2 * Cfg::kTileBytes, &f1map, bar_bfull + sb, x0 - RAD, 0, y0 - RAD + c, b);
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
      cp_async_commit();
      cp_async_wait<0>();
    }
    conv_bar_sync();                               // every group reads all of g2
    mbar_wait(bar_g, 0);
    const uint32_t xmask = 0x1FFu << tx;           // window columns qx with 0 <= qx - tx <= 8
    for (int c = grp; c < QH; c += kGroups) {
      const int s = c % NS, sb = c % NB;
      if (c >= NS) {
        mbar_wait(bar_free + s, ((c / NS) - 1) & 1);
        tc::fence_after_sync();
      }
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
      mbar_wait(bar_bfull + sb, (c / NB) & 1);
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
    }
    // ---- epilogue: D_r[pixel][channel] -> g{f1,f2}[b][channel][y][x] * (1/C); 16-channel chunks dealt round robin to the groups
    mbar_wait(bar_done, 0);
    tc::fence_after_sync();
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
  if (warp == 1) tc::tmem_dealloc(tmem, Cfg::kTmemCols);
}

template <int C>
int launch_bwd_tc(const CUtensorMap& gt, const CUtensorMap& m1, const CUtensorMap& m2, const float* gout, long long gout_bs,
                  float* gf1, float* gf2, int B, int H, int W, cudaStream_t stream) {
  auto kern = cost_volume_bwd_tc_kernel<C>;
  UOF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<C>::kSmem));
  kern<<<dim3(ceil_div(W, TW), ceil_div(H, TH), B), kTcThreads, TcCfg<C>::kSmem, stream>>>(gt, m1, m2, gout, gout_bs, gf1, gf2,
                                                                                            H, W, 1.0f / (float)C);
  count_launch();
  return check_launch("cost_volume_bwd (tcgen05)");
}

}  // namespace

// Tensor-core backward.  Returns false when it does not apply (C not in {32,64,96,128}, W % 4 != 0, unaligned pointers,
// small levels that do not fill the machine, UOF_CV_NO_TC=1, no driver entry point); otherwise launches and stores the status.
bool bwd_tc(const float* gout, long long gout_bs, const float* f1, const float* f2, float* gf1, float* gf2, int B, int C,
            int H, int W, cudaStream_t stream, int* rc) {
  const bool off = getenv("UOF_CV_NO_TC") != nullptr;          // read per call: tests and benches switch paths at run time
  const bool force = getenv("UOF_CV_FORCE_TC") != nullptr;
  if (off || !(C == 32 || C == 64 || C == 96 || C == 128)) return false;
  if (W % 4 != 0 || gout_bs % 4 != 0 || !aligned16(gout) || !aligned16(f1) || !aligned16(f2)) return false;
  if (B > 65535 || ceil_div(H, TH) > 65535) return false;
  const long long tiles = (long long)ceil_div(W, TW) * ceil_div(H, TH) * B;
  if (!force && tiles < 2 * kNumSMs) return false;          // the latency-bound small levels keep their own kernels
  CUtensorMap gt, m1, m2;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NDISP, (cuuint64_t)B};
    const cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)gout_bs * 4};
    const cuuint32_t box_tile[4] = {TW, TH, NDISP, 1};
    if (!make_map_4d(&gt, gout, dims, str, box_tile, CU_TENSOR_MAP_SWIZZLE_NONE)) return false;
  }
  {
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t str[3] = {(cuuint64_t)W * H * 4, (cuuint64_t)W * 4, (cuuint64_t)W * H * C * 4};
    const cuuint32_t box[4] = {QW, (cuuint32_t)C, 1, 1};
    if (!make_map_4d(&m1, f1, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B) ||
        !make_map_4d(&m2, f2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B))
      return false;
  }
  switch (C) {
    case 32: *rc = launch_bwd_tc<32>(gt, m1, m2, gout, gout_bs, gf1, gf2, B, H, W, stream); break;
    case 64: *rc = launch_bwd_tc<64>(gt, m1, m2, gout, gout_bs, gf1, gf2, B, H, W, stream); break;
    case 96: *rc = launch_bwd_tc<96>(gt, m1, m2, gout, gout_bs, gf1, gf2, B, H, W, stream); break;
    default: *rc = launch_bwd_tc<128>(gt, m1, m2, gout, gout_bs, gf1, gf2, B, H, W, stream); break;
  }
  return true;
}

}  // namespace cv
}  // namespace uof
