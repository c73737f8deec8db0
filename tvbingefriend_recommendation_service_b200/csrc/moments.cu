// Exact float64 moments of the text similarity matrix (and of its products with the genre and
// metadata matrices) WITHOUT forming N x N -- the part of get_similarity_statistics
// (ml/similarity_computer.py:171-190) that the fp16 tensor-core sweep cannot deliver to 1e-6:
//
//   sum_ij t_ij        = || X^T 1 ||^2            column sums            [V]
//   sum_ij t_ij^2      = || X^T X ||_F^2          vocabulary Gram matrix [V, V]  (upper triangle kept)
//   sum_ij g_ij t_ij   = || Ghat^T X ||_F^2       Ghat = genre bits / sqrt(popcount)     [64, V]
//   sum_ij m_ij t_ij   = || Mhat^T X ||_F^2       Mhat = one-hot bits * per-show scale   [32, V]
//
// (t = X X^T, g = Ghat Ghat^T, m = Mhat Mhat^T, all symmetric; sums run over ALL (i, j), the caller
// subtracts the diagonal terms, also returned here, and halves.)  X is the normalised text CSR.
// All accumulation is float64 atomics: exact up to rounding, order-independent to ~1e-15 relative.
#include "internal.cuh"

namespace {

__global__ void __launch_bounds__(256)
moments_scatter_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                       const double* __restrict__ values, const TvbfColSide* __restrict__ cs,
                       const float* __restrict__ meta_scale, int meta_hstack, int n_rows, int vocab,
                       double* __restrict__ colsum, double* __restrict__ mg, double* __restrict__ mm,
                       double* __restrict__ out8) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const int64_t b = indptr[row], e = indptr[row + 1];
  const TvbfColSide c = cs[row];
  const int gn = __popcll(c.genre_bits), mn = __popc(c.meta_bits);
  const double ghat = gn ? 1.0 / sqrt(static_cast<double>(gn)) : 0.0;
  // MEAN3: m_ij = matches / 3 = sum_bits (1/sqrt3)(1/sqrt3); HSTACK: matches / sqrt(n_i n_j)
  const double mhat = meta_hstack ? (mn ? 1.0 / sqrt(static_cast<double>(mn)) : 0.0) : 0.57735026918962576451;
  (void)meta_scale;
  double tii = 0.0;
  for (int64_t i = b + lane; i < e; i += 32) {
    const double x = values[i];
    const int col = indices[i];
    tii += x * x;
    atomicAdd(colsum + col, x);
    unsigned long long gb = c.genre_bits;
    while (gb) {
      const int bit = __ffsll(static_cast<long long>(gb)) - 1;
      gb &= gb - 1;
      atomicAdd(mg + static_cast<size_t>(bit) * vocab + col, x * ghat);
    }
    unsigned int mb = c.meta_bits;
    while (mb) {
      const int bit = __ffs(mb) - 1;
      mb &= mb - 1;
      atomicAdd(mm + static_cast<size_t>(bit) * vocab + col, x * mhat);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tii += __shfl_xor_sync(0xffffffffu, tii, o);
  if (lane == 0) {
    const double gii = gn ? 1.0 : 0.0;   // cosine of a non-empty row with itself
    const double mii = meta_hstack ? (mn ? 1.0 : 0.0) : static_cast<double>(mn) / 3.0;
    atomicAdd(out8 + 4, tii);
    atomicAdd(out8 + 5, tii * tii);
    atomicAdd(out8 + 6, gii * tii);
    atomicAdd(out8 + 7, mii * tii);
  }
}

// upper triangle of the vocabulary Gram matrix: G[c, d] += x_ic x_id for c <= d (indices are sorted
// within a row, so entry pairs (a <= b) give c <= d); one warp per row, lanes stride the pairs
__global__ void __launch_bounds__(256)
gram_scatter_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                    const double* __restrict__ values, int n_rows, int vocab, double* __restrict__ gram) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const int64_t b = indptr[row];
  const int len = static_cast<int>(indptr[row + 1] - b);
  for (int a = 0; a < len; ++a) {
    const double xa = values[b + a];
    double* grow = gram + static_cast<size_t>(indices[b + a]) * vocab;
    for (int q = a + lane; q < len; q += 32) atomicAdd(grow + indices[b + q], xa * values[b + q]);
  }
}

// out += sum of squares of `count` doubles; `gram` mode: the array is a [vocab, vocab] upper triangle,
// off-diagonal entries count twice
__global__ void __launch_bounds__(256)
sum_squares_kernel(const double* __restrict__ x, size_t count, int vocab, int gram, double* __restrict__ out) {
  double s = 0.0;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    const double v = x[i];
    double w = 1.0;
    if (gram) {
      const size_t r = i / vocab, c = i - r * vocab;
      w = c > r ? 2.0 : (c == r ? 1.0 : 0.0);
    }
    s += w * v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(out, t);
  }
}

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

extern "C" {

size_t tvbf_text_moments_workspace_bytes(const tvbf_features* f, int32_t with_gram) {
  if (f == nullptr || f->vocab <= 0) return 0;
  const size_t v = static_cast<size_t>(f->vocab);
  size_t bytes = align256(v * 8) + align256(64 * v * 8) + align256(32 * v * 8);
  if (with_gram) bytes += align256(v * v * 8);
  return bytes;
}

int tvbf_text_moments(const tvbf_features* f, int32_t with_gram, double* out8, void* workspace,
                      size_t workspace_bytes, void* stream) {
  TVBF_REQUIRE(f && out8 && workspace, "tvbf_text_moments: NULL argument");
  TVBF_REQUIRE(f->genre_mode != TVBF_GROUP_FOLDED && f->meta_mode != TVBF_GROUP_FOLDED,
               "tvbf_text_moments needs binary genre / one-hot metadata features");
  const size_t need = tvbf_text_moments_workspace_bytes(f, with_gram);
  if (workspace_bytes < need) {
    tvbf_set_error("tvbf_text_moments: workspace too small: %zu < %zu", workspace_bytes, need);
    return TVBF_ERR_WORKSPACE;
  }
  auto st = static_cast<cudaStream_t>(stream);
  const int n = f->n_shows, v = f->vocab;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  double* colsum = reinterpret_cast<double*>(ws);
  double* mg = reinterpret_cast<double*>(ws + align256(static_cast<size_t>(v) * 8));
  double* mm = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(mg) + align256(static_cast<size_t>(64) * v * 8));
  double* gram = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(mm) + align256(static_cast<size_t>(32) * v * 8));
  TVBF_CUDA_OK(cudaMemsetAsync(workspace, 0, need, st));
  TVBF_CUDA_OK(cudaMemsetAsync(out8, 0, 8 * sizeof(double), st));
  const unsigned row_blocks = static_cast<unsigned>((static_cast<size_t>(n) * 32 + 255) / 256);
  moments_scatter_kernel<<<row_blocks, 256, 0, st>>>(f->text_indptr, f->text_indices, f->text_values,
                                                     static_cast<const TvbfColSide*>(f->col_side), f->meta_scale,
                                                     f->meta_kind == TVBF_META_HSTACK ? 1 : 0, n, v, colsum, mg, mm,
                                                     out8);
  TVBF_LAUNCH_OK("moments_scatter_kernel");
  sum_squares_kernel<<<148, 256, 0, st>>>(colsum, static_cast<size_t>(v), v, 0, out8 + 0);
  TVBF_LAUNCH_OK("sum_squares_kernel");
  sum_squares_kernel<<<148, 256, 0, st>>>(mg, static_cast<size_t>(64) * v, v, 0, out8 + 2);
  TVBF_LAUNCH_OK("sum_squares_kernel");
  sum_squares_kernel<<<148, 256, 0, st>>>(mm, static_cast<size_t>(32) * v, v, 0, out8 + 3);
  TVBF_LAUNCH_OK("sum_squares_kernel");
  if (with_gram) {
    gram_scatter_kernel<<<row_blocks, 256, 0, st>>>(f->text_indptr, f->text_indices, f->text_values, n, v, gram);
    TVBF_LAUNCH_OK("gram_scatter_kernel");
    sum_squares_kernel<<<148 * 8, 256, 0, st>>>(gram, static_cast<size_t>(v) * v, v, 1, out8 + 1);
    TVBF_LAUNCH_OK("sum_squares_kernel");
  }
  return TVBF_OK;
}

}  // extern "C"
