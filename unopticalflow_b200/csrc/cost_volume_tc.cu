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
constexpr int kTcThreads = 192;                 // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: converters / epilogue
constexpr int kGBytes = NDISP * TH * TW * 4;    // one gout tile: 41 472 B

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int C>
struct TcCfg {
  static constexpr int kStages = C <= 32 ? 3 : 4;
  static constexpr int kTileBytes = C * QW * 4;                        // one [C][16] operand tile (C * 64 B, multiple of 512)
  static constexpr int kStageBytes = 4 * kTileBytes;                   // role x {hi, lo}
  static constexpr int kACols = 64;                                    // per stage: role x {hi, lo} x 16 columns
  static constexpr int kTmemCols = 2 * C + kStages * kACols <= 256 ? 256 : 512;
  static constexpr size_t kSmem = 1024 + 2 * (size_t)kGBytes + (size_t)kStages * kStageBytes + 256;
  static_assert(C % 16 == 0 && C >= 16 && C <= 128, "N = C must be a multiple of 16 (M = 128) and fit TMEM");
  static_assert(2 * C + kStages * kACols <= 512, "TMEM budget");
};

template <int C>
__global__ void __launch_bounds__(kTcThreads, C <= 32 ? 2 : 1)
cost_volume_bwd_tc_kernel(const __grid_constant__ CUtensorMap gmap_tile, const __grid_constant__ CUtensorMap f1map,
                          const __grid_constant__ CUtensorMap f2map, const float* __restrict__ gout, long long gout_bs,
                          float* __restrict__ gf1, float* __restrict__ gf2, int H, int W, float inv_c) {
  using Cfg = TcCfg<C>;
  constexpr int S = Cfg::kStages;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte aligned carve-up (the operand tiles need 512 for SWIZZLE_64B)
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* g1 = reinterpret_cast<float*>(base);                          // [81][128]  gout tile
  float* g2 = g1 + NDISP * TH * TW;                                    // [81][128]  sheared gout tile
  unsigned char* ring = base + 2 * kGBytes;                            // [S][role][hi/lo][C][16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + S * Cfg::kStageBytes);
  uint64_t* bar_g = bars;                 // gout tiles landed
  uint64_t* bar_bfull = bars + 1;         // [S] feature rows landed
  uint64_t* bar_ready = bars + 1 + S;     // [S] converters done (A in TMEM, B_lo in smem)
  uint64_t* bar_free = bars + 1 + 2 * S;  // [S] MMAs of the stage retired
  uint64_t* bar_done = bars + 1 + 3 * S;  // all MMAs retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 3 * S);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&gmap_tile);
    tma_prefetch_desc(&f1map);
    tma_prefetch_desc(&f2map);
    mbar_init(bar_g, 1);
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_bfull + s, 1);
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
  // TMEM columns: [0,C) D_0 (gf1)   [C,2C) D_1 (gf2)   [2C + 64 s, +64): stage s = {A_0 hi, A_0 lo, A_1 hi, A_1 lo} x 16
  const uint32_t tmem_a = tmem + 2 * C;

  if (warp == 0) {
    // ================================================= TMA producer =================================================
    if (lane == 0) {
      mbar_expect_tx(bar_g, kGBytes);
      tma_load_4d(g1, &gmap_tile, bar_g, x0, y0, 0, b);
      for (int c = 0; c < QH; ++c) {
        const int s = c % S;
        if (c >= S) mbar_wait(bar_free + s, ((c / S) - 1) & 1);
        unsigned char* st = ring + s * Cfg::kStageBytes;
        mbar_expect_tx(bar_bfull + s, 2 * Cfg::kTileBytes);
        tma_load_4d(st, &f2map, bar_bfull + s, x0 - RAD, 0, y0 - RAD + c, b);                          // role 0: f2 row
        tma_load_4d(st + 2 * Cfg::kTileBytes, &f1map, bar_bfull + s, x0 - RAD, 0, y0 - RAD + c, b);    // role 1: f1 row
      }
    }
  } else if (warp == 1) {
    // ================================================== MMA issuer ==================================================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::idesc_tf32(128, C);
      for (int c = 0; c < QH; ++c) {
        const int s = c % S;
        const uint32_t par = (c / S) & 1;
        mbar_wait(bar_bfull + s, par);
        mbar_wait(bar_ready + s, par);
        tc::fence_after_sync();
        const uint32_t st = smem_u32(ring + s * Cfg::kStageBytes);
        const uint32_t ta = tmem_a + s * Cfg::kACols;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t d = tmem + r * C;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t bh = tc::smem_desc_kmajor(st + (2 * r) * Cfg::kTileBytes + 32 * ks, 512, tc::kLayoutSw64);
            const uint64_t bl = tc::smem_desc_kmajor(st + (2 * r + 1) * Cfg::kTileBytes + 32 * ks, 512, tc::kLayoutSw64);
            const uint32_t ah = ta + 32 * r + 8 * ks, al = ah + 16;
            tc::mma_tf32_ts(d, ah, bh, idesc, c > 0 || ks > 0);      // hi * hi
            tc::mma_tf32_ts(d, al, bh, idesc, true);                 // lo * hi
            tc::mma_tf32_ts(d, ah, bl, idesc, true);                 // hi * lo
          }
        }
        tc::mma_commit(bar_free + s);      // stage reusable when these MMAs have read their operands
      }
      tc::mma_commit(bar_done);
    }
  } else {
    // ========================================= converters (4 warps) + epilogue =========================================
    const int q = warp & 3;                        // TMEM lane quadrant this warp may touch
    const int m = q * 32 + lane;                   // tile pixel = TMEM lane
    const int ty = m >> 3, tx = m & 7;
    const int ty_lo = q * 4, ty_hi = q * 4 + 3;    // rows of this warp
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int ct = threadIdx.x - 64;               // 0..127: converter thread index for the B pass
    {
      // sheared gout column of this lane: g2[d][m] = gout[b, d, y - (i-4), x - (j-4)], zero outside the image.  Thread-private
      // (lane m only ever reads column m), so the copies need no barrier, just this thread's own cp.async wait.
      const int y = y0 + ty, x = x0 + tx;
      const float* gb = gout + (size_t)b * gout_bs;
      const size_t plane = (size_t)H * W;
#pragma unroll
      for (int i = 0; i < ND; ++i) {
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
    mbar_wait(bar_g, 0);
    for (int c = 0; c < QH; ++c) {
      const int s = c % S;
      if (c >= S) {
        mbar_wait(bar_free + s, ((c / S) - 1) & 1);
        tc::fence_after_sync();
      }
      // ---- A chunks of window row qy = c -> TMEM
      const uint32_t ta = tmem_a + s * Cfg::kACols + lane_addr;
      uint32_t h0[16], l0[16], h1[16], l1[16];
      if (c >= ty_lo && c <= ty_hi + 2 * RAD) {    // warp-uniform: some lane of the warp is inside the band
        const int dyi = c - ty;
        const bool band_y = dyi >= 0 && dyi <= 2 * RAD;
        const int d_base = dyi * ND - tx;          // d1 = d_base + qx
        const float* p0 = g1 + d_base * (TH * TW) + m;
        const float* p1 = g2 + (NDISP - 1 - d_base) * (TH * TW) + m;
#pragma unroll
        for (int qx = 0; qx < QW; ++qx) {
          const bool in = band_y && qx >= tx && qx <= tx + 2 * RAD;
          const float v0 = in ? p0[qx * (TH * TW)] : 0.0f;
          const float v1 = in ? p1[-qx * (TH * TW)] : 0.0f;
          h0[qx] = tc::tf32_hi(v0);
          l0[qx] = tc::tf32_lo(v0, h0[qx]);
          h1[qx] = tc::tf32_hi(v1);
          l1[qx] = tc::tf32_lo(v1, h1[qx]);
        }
      } else {
#pragma unroll
        for (int qx = 0; qx < QW; ++qx) h0[qx] = l0[qx] = h1[qx] = l1[qx] = 0u;
      }
      tc::tmem_st16(ta, h0);
      tc::tmem_st16(ta + 16, l0);
      tc::tmem_st16(ta + 32, h1);
      tc::tmem_st16(ta + 48, l1);
      // ---- B_lo = B - trunc_tf32(B) for both feature rows (same swizzled position in the lo buffer)
      mbar_wait(bar_bfull + s, (c / S) & 1);
      {
        float4* st = reinterpret_cast<float4*>(ring + s * Cfg::kStageBytes);
        constexpr int kVec = Cfg::kTileBytes / 16;                     // float4 per tile = 4 C
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float4* src = st + (2 * r) * kVec;
          float4* dst = st + (2 * r + 1) * kVec;
#pragma unroll
          for (int i = ct; i < kVec; i += 128) {
            const float4 v = src[i];
            float4 o;
            o.x = __uint_as_float(tc::tf32_lo(v.x, tc::tf32_hi(v.x)));
            o.y = __uint_as_float(tc::tf32_lo(v.y, tc::tf32_hi(v.y)));
            o.z = __uint_as_float(tc::tf32_lo(v.z, tc::tf32_hi(v.z)));
            o.w = __uint_as_float(tc::tf32_lo(v.w, tc::tf32_hi(v.w)));
            dst[i] = o;
          }
        }
      }
      tc::fence_proxy_async();
      tc::wait_st();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ready + s);
    }
    // ---- epilogue: D_r[pixel][channel] -> g{f1,f2}[b][channel][y][x] * (1/C)
    mbar_wait(bar_done, 0);
    tc::fence_after_sync();
    const int y = y0 + ty, x = x0 + tx;
    const bool live = y < H && x < W;
    const size_t plane = (size_t)H * W;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float* dst = (r == 0 ? gf1 : gf2) + (size_t)b * C * plane + (size_t)y * W + x;
#pragma unroll
      for (int j = 0; j < C / 16; ++j) {
        uint32_t v[16];
        tc::tmem_ld16(tmem + lane_addr + r * C + 16 * j, v);
        tc::wait_ld();
        if (live) {
#pragma unroll
          for (int k = 0; k < 16; ++k) dst[(size_t)(16 * j + k) * plane] = __uint_as_float(v[k]) * inv_c;
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
