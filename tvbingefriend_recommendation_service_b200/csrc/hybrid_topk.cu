// K1 + K4 -- the hot kernel: TMA-fed tcgen05 text GEMM (X * X^T, fp16/bf16 operands, fp32
// accumulators in TMEM) whose epilogue adds the popcount genre score and the category-id
// metadata score with the reference's weights and keeps a running per-row candidate list, so
// the N x N matrix is never written anywhere.
//
// Replaces, for all source rows at once, the body of the reference's hot loop
// (scripts/populate_database.py:170-195): five cosine_similarity calls, the (p+t+l)/3 metadata
// mean, the weighted sum and the full argsort.  The selection made here is a conservative
// CANDIDATE selection on an upper bound U >= exact score; rescore.cu recomputes the candidates in
// fp64 and certifies (or repairs) each row, so the final table does not depend on fp16 rounding.
//
// Kernel anatomy (one CTA per SM, persistent over work items = (row block of 128, column split)):
//   warp 0   lane 0 : TMA producer  -- A (128 x 64) and B (256 x 64) operand tiles, 4-stage ring,
//                                      plus the column-side records of each 256-column tile
//   warp 1   lane 0 : MMA issuer    -- tcgen05.mma cta_group::1 kind::f16, M=128 N=256 K=16,
//                                      double-buffered 128x256 fp32 accumulators (2 x 256 TMEM cols)
//   warps 2-5       : epilogue      -- tcgen05.ld 32x32b: thread t owns accumulator row t; scores,
//                                      thresholds and appends to its row's candidate list
#include "internal.cuh"

#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through the runtime

namespace tvbf {

constexpr int BM = 128;     // rows per CTA tile (= TMEM lanes)
constexpr int BN = 256;     // columns per tile (= TMEM columns per accumulator)
constexpr int BK = 64;      // K elements per stage: 64 halves = one 128-byte swizzle row
constexpr int STAGES = 4;
constexpr int UMMA_K = 16;
constexpr uint32_t A_BYTES = BM * BK * 2;
constexpr uint32_t B_BYTES = BN * BK * 2;
constexpr uint32_t COL_BYTES = BN * sizeof(TvbfColSide);
constexpr uint32_t MS_BYTES = BN * sizeof(float);
constexpr uint32_t OFF_A = 0;
constexpr uint32_t OFF_B = OFF_A + STAGES * A_BYTES;
constexpr uint32_t OFF_COL = OFF_B + STAGES * B_BYTES;
constexpr uint32_t OFF_MS = OFF_COL + 2 * COL_BYTES;
constexpr uint32_t OFF_BAR = OFF_MS + 2 * MS_BYTES;
constexpr int NUM_BARS = 2 * STAGES + 6;
constexpr uint32_t OFF_TMEM = OFF_BAR + NUM_BARS * 8;
constexpr uint32_t SMEM_USED = OFF_TMEM + 16;
constexpr uint32_t SMEM_BYTES = SMEM_USED + 1024;  // slack for manual 1024-byte alignment
constexpr int NUM_THREADS = 192;
constexpr int EPI_WARP0 = 2;


// ---- warp-cooperative bitonic sort (descending) of 32*E 64-bit keys, E per lane ----------------
// element index = q * 32 + lane
template <int E>
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long (&key)[E], int lane) {
  constexpr int NTOT = 32 * E;
#pragma unroll
  for (int k = 2; k <= NTOT; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int dq = j >> 5;
#pragma unroll
        for (int q = 0; q < E; ++q) {
          const int q2 = q ^ dq;
          if (q2 > q) {
            const bool desc = (((q * 32) & k) == 0);  // lane bits are below k's bit here
            unsigned long long a = key[q], b = key[q2];
            const bool sw = desc ? (a < b) : (a > b);
            key[q] = sw ? b : a;
            key[q2] = sw ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < E; ++q) {
          const unsigned long long other = __shfl_xor_sync(kFullMask, key[q], j);
          const int idx = q * 32 + lane;
          const bool lower = (lane & j) == 0;
          const bool desc = ((idx & k) == 0);
          const unsigned long long mx = key[q] > other ? key[q] : other;
          const unsigned long long mn = key[q] > other ? other : key[q];
          key[q] = (lower == desc) ? mx : mn;
        }
      }
    }
  }
}

// Sort the `n` entries of one row's list and write the best min(n, kp) to dst (may alias src).
// Returns (to every lane) the score of entry kp-1, i.e. the bound on everything dropped.
template <int E>
__device__ __forceinline__ float warp_compact(const uint2* src, int n, uint2* dst, int kp,
                                              int lane) {
  unsigned long long key[E];
  __syncwarp();  // the owner lane's appends (st.cg) are ordered before these loads
#pragma unroll
  for (int q = 0; q < E; ++q) {
    const int idx = q * 32 + lane;
    key[q] = 0ull;
    if (idx < n) {
      const uint2 e = __ldcg(src + idx);
      key[q] = (static_cast<unsigned long long>(f32_orderable(__uint_as_float(e.x))) << 32) |
               static_cast<unsigned long long>(0xFFFFFFFFu - e.y);
    }
  }
  __syncwarp();  // all loads done before anyone overwrites (dst may alias src)
  bitonic_sort_desc<E>(key, lane);
  const int keep = n < kp ? n : kp;
  float bound = 0.0f;
#pragma unroll
  for (int q = 0; q < E; ++q) {
    const int idx = q * 32 + lane;
    const float sc = f32_from_orderable(static_cast<uint32_t>(key[q] >> 32));
    if (idx < keep)
      __stcg(dst + idx, make_uint2(__float_as_uint(sc), 0xFFFFFFFFu - static_cast<uint32_t>(key[q])));
    // entry kp-1 lives in register (kp-1)/32 of lane (kp-1)%32
    const float cand = __shfl_sync(kFullMask, sc, (kp - 1) & 31);
    if (q == ((kp - 1) >> 5)) bound = cand;
  }
  __syncwarp();
  return bound;
}

struct ItemCoord {
  int rb;      // row block within the shard, -1 = nothing to do
  int split;
  int tile0, tile1;
};

__device__ __forceinline__ ItemCoord item_coord(const K1Params& p, int item) {
  const int per_group = p.rb_per_group * p.splits;
  const int g = item / per_group;
  const int w = item - g * per_group;
  ItemCoord c;
  c.split = w / p.rb_per_group;
  c.rb = g * p.rb_per_group + (w - c.split * p.rb_per_group);
  if (c.rb >= p.rb_count) c.rb = -1;
  c.tile0 = static_cast<int>(static_cast<long long>(p.col_tiles) * c.split / p.splits);
  c.tile1 = static_cast<int>(static_cast<long long>(p.col_tiles) * (c.split + 1) / p.splits);
  return c;
}

template <int E, bool kDump>
__global__ void __launch_bounds__(NUM_THREADS, 1)
hybrid_topk_kernel(const __grid_constant__ CUtensorMap tmap_a,
                   const __grid_constant__ CUtensorMap tmap_b, const K1Params p,
                   const uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* full = bars;                      // [STAGES]
  uint64_t* empty = bars + STAGES;            // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;     // [2]
  uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2]
  uint64_t* col_full = bars + 2 * STAGES + 4;   // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 128);
      mbar_init(&col_full[b], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int n_items = kDump ? 1
                            : ((p.rb_count + p.rb_per_group - 1) / p.rb_per_group) *
                                  p.rb_per_group * p.splits;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        ItemCoord c = item_coord(p, item);
        if (kDump) { c.rb = 0; c.tile0 = 0; c.tile1 = 1; }
        if (c.rb < 0) continue;
        const int row0 = kDump ? p.row_begin : p.row_begin + c.rb * BM;
        for (int jt = c.tile0; jt < c.tile1; ++jt, ++it) {
          const uint32_t b = it & 1;
          const int col0 = kDump ? p.dump_col0 : jt * BN;
          mbar_wait(&acc_empty[b], ((it >> 1) & 1) ^ 1);  // column-side buffer b is free
          if (!kDump) {
            mbar_arrive_expect_tx(&col_full[b], COL_BYTES + MS_BYTES);
            bulk_load_1d(smem + OFF_COL + b * COL_BYTES, p.col_side + col0, COL_BYTES,
                         &col_full[b]);
            bulk_load_1d(smem + OFF_MS + b * MS_BYTES, p.meta_scale + col0, MS_BYTES,
                         &col_full[b]);
          }
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], A_BYTES + B_BYTES);
            tma_load_2d(smem + OFF_A + stage * A_BYTES, &tmap_a, &full[stage], kb * BK, row0);
            tma_load_2d(smem + OFF_B + stage * B_BYTES, &tmap_b, &full[stage], kb * BK, col0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        ItemCoord c = item_coord(p, item);
        if (kDump) { c.rb = 0; c.tile0 = 0; c.tile1 = 1; }
        if (c.rb < 0) continue;
        for (int jt = c.tile0; jt < c.tile1; ++jt, ++it) {
          const uint32_t b = it & 1;
          mbar_wait(&acc_empty[b], ((it >> 1) & 1) ^ 1);  // epilogue drained accumulator b
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + b * BN;
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint64_t da = umma_desc_sw128(smem_u32(smem + OFF_A + stage * A_BYTES));
            const uint64_t db = umma_desc_sw128(smem_u32(smem + OFF_B + stage * B_BYTES));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 32 bytes (16 halves) along K inside the 128-byte swizzle row
              umma_f16(tmem_d, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2),
                       idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty[stage]);  // smem slot reusable once these MMAs have read it
            if (kb == p.k_blocks - 1) umma_commit(&acc_full[b]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    // =============================== epilogue ===================================
    const int quarter = warp & 3;                 // TMEM lanes this warp may touch
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t tmem_lane = static_cast<uint32_t>(quarter * 32) << 16;
    constexpr int CAP = 32 * E;
    uint2* my_list = p.scratch + (static_cast<size_t>(blockIdx.x) * BM + row_in_tile) * CAP;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      ItemCoord c = item_coord(p, item);
      if (kDump) { c.rb = 0; c.tile0 = 0; c.tile1 = 1; }
      if (c.rb < 0) continue;
      const int row = (kDump ? p.row_begin : p.row_begin + c.rb * BM) + row_in_tile;
      const bool row_valid = row < p.row_end;
      // row-side operands of the fused scores
      unsigned long long g_bits = 0ull;
      float rn_wg = 0.0f, ci_wm8 = 0.0f;
      uint32_t ids_row = 0xFEFEFEFEu;
      if (row_valid && !kDump) {
        const TvbfColSide rs = p.col_side[row];
        g_bits = rs.genre_bits;
        rn_wg = rs.genre_rnorm * p.w_genre;
        // "none" (0xFF) must never equal a column's "none": recode it to 0xFE on the row side
        const uint32_t none = __vcmpeq4(rs.meta_ids, 0xFFFFFFFFu);
        ids_row = (rs.meta_ids & ~none) | (0xFEFEFEFEu & none);
        ci_wm8 = p.meta_scale[row] * p.w_meta8;
      }
      float theta = row_valid ? p.theta_init : __int_as_float(0x7f800000);  // +inf: never append
      int cnt = 0;
      bool dropped = false;
      const int self_col = p.exclude_self ? row : -1;

      for (int jt = c.tile0; jt < c.tile1; ++jt, ++it) {
        const uint32_t b = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int col0 = kDump ? p.dump_col0 : jt * BN;
        if (!kDump) mbar_wait(&col_full[b], ph);
        mbar_wait(&acc_full[b], ph);
        tc_fence_after();
        const TvbfColSide* scol = reinterpret_cast<const TvbfColSide*>(smem + OFF_COL + b * COL_BYTES);
        const float* sms = reinterpret_cast<const float*>(smem + OFF_MS + b * MS_BYTES);
        const uint32_t taddr = tmem_base + tmem_lane + b * BN;

        uint32_t acc[2][32];
        tmem_ld_32x32(taddr, acc[0]);
#pragma unroll
        for (int ch = 0; ch < BN / 32; ++ch) {
          tmem_ld_wait();
          if (ch + 1 < BN / 32) tmem_ld_32x32(taddr + (ch + 1) * 32, acc[(ch + 1) & 1]);
          if (kDump) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              p.dump[row_in_tile * BN + ch * 32 + e] = __uint_as_float(acc[ch & 1][e]);
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float a = __uint_as_float(acc[ch & 1][e]);
              const TvbfColSide cs = scol[ch * 32 + e];
              const float ms = sms[ch * 32 + e];
              const float gdot = static_cast<float>(__popcll(g_bits & cs.genre_bits)) * cs.genre_rnorm;
              const float mdot = static_cast<float>(__popc(__vcmpeq4(ids_row, cs.meta_ids))) * ms;
              float u = fmaf(gdot, rn_wg, fmaf(mdot, ci_wm8, p.eps));
              u = fmaf(a, p.w_text, u);
              u = fmaf(fabsf(a), p.w_text_err, u);
              if (u > theta) {
                const int col = col0 + ch * 32 + e;
                if (col != self_col && col < p.n_shows) {
                  __stcg(my_list + cnt, make_uint2(__float_as_uint(u), static_cast<uint32_t>(col)));
                  ++cnt;
                }
              }
            }
            // keep 32 free slots for the next chunk; compact rows that are nearly full
            unsigned need = __ballot_sync(kFullMask, cnt > CAP - 32);
            while (need) {
              const int src_lane = __ffs(need) - 1;
              need &= need - 1;
              const uint2* lp = reinterpret_cast<const uint2*>(__shfl_sync(
                  kFullMask, reinterpret_cast<unsigned long long>(my_list), src_lane));
              const int n = __shfl_sync(kFullMask, cnt, src_lane);
              const float bound = warp_compact<E>(lp, n, const_cast<uint2*>(lp), p.kp, lane);
              if (lane == src_lane) {
                cnt = p.kp;
                theta = fmaxf(theta, bound);
                dropped = true;
              }
            }
          }
        }
        // accumulator b and column-side buffer b are free again
        tc_fence_before();
        mbar_arrive(&acc_empty[b]);
      }

      if (!kDump) {
        // final compaction of every row of this warp: sorted best-kp list -> cand
        const int rows_in_shard = p.row_end - p.row_begin;
        for (int src_lane = 0; src_lane < 32; ++src_lane) {
          const int r = c.rb * BM + quarter * 32 + src_lane;  // row within the shard
          if (r >= rows_in_shard) break;                      // warp-uniform
          const uint2* lp = reinterpret_cast<const uint2*>(__shfl_sync(
              kFullMask, reinterpret_cast<unsigned long long>(my_list), src_lane));
          const int n = __shfl_sync(kFullMask, cnt, src_lane);
          uint2* dst = p.cand + (static_cast<size_t>(r) * p.splits + c.split) * p.kp;
          const float bound = warp_compact<E>(lp, n, dst, p.kp, lane);
          if (lane == src_lane) {
            float th = dropped ? theta : __int_as_float(0xff800000);  // -inf: nothing dropped
            if (n > p.kp) th = fmaxf(dropped ? theta : bound, bound);
            p.cand_cnt[static_cast<size_t>(r) * p.splits + c.split] = n < p.kp ? n : p.kp;
            p.cand_theta[static_cast<size_t>(r) * p.splits + c.split] = th;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encoder(EncodeTiledFn* fn) {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  TVBF_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
  if (sym == nullptr || qres != cudaDriverEntryPointSuccess) {
    tvbf_set_error("cuTensorMapEncodeTiled is not available from this driver");
    return TVBF_ERR_CUDA;
  }
  *fn = reinterpret_cast<EncodeTiledFn>(sym);
  return TVBF_OK;
}

// 2-D map over the [n_pad, k_pad] 16-bit operand; box = 64 K-elements x box_rows, 128B swizzle.
static int make_operand_map(const tvbf_features* f, int box_rows, CUtensorMap* out) {
  EncodeTiledFn enc = nullptr;
  int rc = get_encoder(&enc);
  if (rc != TVBF_OK) return rc;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(f->k_pad), static_cast<cuuint64_t>(f->n_pad)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(f->k_pad) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = f->text_dtype == TVBF_TEXT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                          : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(out, dt, 2, const_cast<void*>(f->operand), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tvbf_set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return TVBF_ERR_CUDA;
  }
  return TVBF_OK;
}

int k1_entries_per_lane(int k) {
  if (k <= 48) return 4;
  if (k <= 160) return 8;
  if (k <= 400) return 16;
  return 0;
}

int k1_default_candidates(int k) {
  const int e = k1_entries_per_lane(k);
  const int cap = 32 * e;
  int kp = k <= 48 ? (k + 12 > 32 ? k + 12 : 32) : k + 28;
  kp = (kp + 7) & ~7;
  if (kp > cap - 64) kp = cap - 64;
  return kp;
}

int k1_choose_splits(int rb_count, int col_tiles, int sm_count) {
  // concurrently running CTAs should cover (row blocks) x (column splits) so that both operand
  // streams are shared through L2; prefer the split count with the best wave efficiency.
  int best = 1;
  double best_eff = -1.0;
  for (int s = 1; s <= 8; ++s) {
    if (s > 1 && col_tiles / s < 4) break;
    const int r = sm_count / s;
    if (r < 1) break;
    const int groups = (rb_count + r - 1) / r;
    double eff = static_cast<double>(rb_count) / (static_cast<double>(groups) * r) *
                 (static_cast<double>(r * s) / sm_count);
    // mild preference for 2..4 splits: fewer lists to merge, both operands still L2-shared
    if (s >= 2 && s <= 4) eff += 0.01;
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
  }
  return best;
}

template <int E, bool kDump>
static int launch_k1(const tvbf_features* f, const K1Params& kp, int grid, cudaStream_t st) {
  CUtensorMap ta, tb;
  int rc = make_operand_map(f, BM, &ta);
  if (rc != TVBF_OK) return rc;
  rc = make_operand_map(f, BN, &tb);
  if (rc != TVBF_OK) return rc;
  auto kern = hybrid_topk_kernel<E, kDump>;
  TVBF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(SMEM_BYTES)));
  const uint32_t idesc = umma_idesc_f16(f->text_dtype == TVBF_TEXT_BF16 ? 1u : 0u, BM, BN);
  kern<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ta, tb, kp, idesc);
  TVBF_LAUNCH_OK("hybrid_topk_kernel");
  return TVBF_OK;
}

int k1_launch(const tvbf_features* f, const K1Params& kp, int entries_per_lane, int grid,
              cudaStream_t st) {
  switch (entries_per_lane) {
    case 4: return launch_k1<4, false>(f, kp, grid, st);
    case 8: return launch_k1<8, false>(f, kp, grid, st);
    case 16: return launch_k1<16, false>(f, kp, grid, st);
    default:
      tvbf_set_error("unsupported candidate capacity (entries per lane %d)", entries_per_lane);
      return TVBF_ERR_INVALID;
  }
}

int k1_launch_dump(const tvbf_features* f, const K1Params& kp, cudaStream_t st) {
  return launch_k1<4, true>(f, kp, 1, st);
}

}  // namespace tvbf
