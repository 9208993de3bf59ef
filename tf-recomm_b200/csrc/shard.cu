// Row-sharded tables (BASELINE configs[4]): rows of each table and their Adam state live on rank
// (id mod G); this is the owner's half of the per-step row exchange.  The reference has no distributed code at
// all (single tf.Session, svd_train_val.py:55) -- this is the B200-native equivalent north_star asks for.
#include "common.cuh"

namespace tfr {

template <int VEC, int L>
__global__ void __launch_bounds__(256) shard_gather_rows_kernel(const float* __restrict__ feat,
                                                                const float* __restrict__ bias, int rows_local,
                                                                int dim, int64_t fstride,
                                                                const int32_t* __restrict__ ids, int64_t B,
                                                                int n_ranks, int rank, float* __restrict__ out_feat,
                                                                float* __restrict__ out_bias,
                                                                int32_t* __restrict__ keys) {
  const int lane = threadIdx.x & (L - 1);
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  if (b >= B) return;
  const int32_t id = ids[b];
  const bool mine = (id % n_ranks) == rank;
  const int32_t local = id / n_ranks;
  const int n_units = dim / VEC;
  for (int unit = lane; unit < n_units; unit += L) {
    if constexpr (VEC == 4) {
      float4 x = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (mine) x = ld_gather_f4(reinterpret_cast<const float4*>(feat + (size_t)local * fstride) + unit);
      reinterpret_cast<float4*>(out_feat + (size_t)b * dim)[unit] = x;
    } else {
      out_feat[(size_t)b * dim + unit] = mine ? ld_gather_f1(feat + (size_t)local * fstride + unit) : 0.0f;
    }
  }
  if (lane == 0) {
    out_bias[b] = mine ? bias[local] : 0.0f;
    keys[b] = mine ? local : rows_local;
  }
}

}  // namespace tfr

using namespace tfr;

extern "C" int tfr_shard_gather_rows(const float* feat_local, const float* bias_local, int64_t rows_local, int32_t dim,
                                     int64_t feat_stride,
                                     const int32_t* ids, int64_t B, int32_t n_ranks, int32_t rank, float* out_feat,
                                     float* out_bias, int32_t* local_keys, void* stream) {
  TFR_CHECK_ARG(B >= 0 && dim > 0 && n_ranks >= 1 && rank >= 0 && rank < n_ranks && rows_local >= 0 &&
                rows_local < ((int64_t)1 << 31));
  if (B == 0) return TFR_OK;
  TFR_CHECK_ARG(feat_local && bias_local && ids && out_feat && out_bias && local_keys);
  const RowGeom g = row_geom(dim);
  const int groups_per_cta = 256 / g.lanes;
  const unsigned grid = (unsigned)((B + groups_per_cta - 1) / groups_per_cta);
  cudaStream_t st = (cudaStream_t)stream;
#define TFR_SG_CASE(V, LL)                                                                                        \
  if (g.vec == V && g.lanes == LL) {                                                                              \
    TFR_PREP((shard_gather_rows_kernel<V, LL>));                                                                  \
    shard_gather_rows_kernel<V, LL><<<grid, 256, 0, st>>>(feat_local, bias_local, (int)rows_local, dim,              \
                                                          feat_stride ? feat_stride : (int64_t)dim, ids, B,   \
                                                          n_ranks, rank, out_feat, out_bias, local_keys);         \
    TFR_LAUNCH_CHECK();                                                                                            \
    return TFR_OK;                                                                                                 \
  }
  TFR_SG_CASE(4, 1) TFR_SG_CASE(4, 2) TFR_SG_CASE(4, 4) TFR_SG_CASE(4, 8) TFR_SG_CASE(4, 16) TFR_SG_CASE(4, 32)
  TFR_SG_CASE(1, 1) TFR_SG_CASE(1, 2) TFR_SG_CASE(1, 4) TFR_SG_CASE(1, 8) TFR_SG_CASE(1, 16) TFR_SG_CASE(1, 32)
#undef TFR_SG_CASE
  set_error("unsupported dim %d", dim);
  return TFR_ERR_INVALID;
}
