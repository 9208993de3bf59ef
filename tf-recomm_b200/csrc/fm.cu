// K5 (forward): 2nd-order factorization machine on CSR rows.
// Replaces forward.py:21-22  fma(x) = mu + x.W + 0.5*(||xV||^2 - sum_f sum_i x_i^2 V_if^2)
// (forward.py writes x.dot(V**2), valid for the 0/1 features fm.py:61-93 builds; the canonical x_i^2
// form is computed, they coincide on one-hot / multi-hot rows -- SURVEY 8a row a19).
// One lane group per CSR row; a lane keeps s_f and q_f for its VEC-wide slice of the factors.
#include "common.cuh"

namespace tfr {

template <int VEC, int L, int UNITS>
__global__ void __launch_bounds__(256) fm_forward_kernel(int64_t n_rows, const int64_t* __restrict__ indptr,
                                                         const int32_t* __restrict__ indices,
                                                         const float* __restrict__ data, const float* __restrict__ w0,
                                                         const float* __restrict__ W, const float* __restrict__ V,
                                                         int dim, float* __restrict__ yhat, float* __restrict__ sums) {
  const int lane = threadIdx.x & (L - 1);
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  if (r >= n_rows) return;
  const int n_units = dim / VEC;
  float s[UNITS][VEC], q[UNITS][VEC];
#pragma unroll
  for (int u = 0; u < UNITS; ++u)
#pragma unroll
    for (int c = 0; c < VEC; ++c) { s[u][c] = 0.0f; q[u][c] = 0.0f; }
  float lin = 0.0f;
  const int64_t p0 = indptr[r], p1 = indptr[r + 1];
  for (int64_t p = p0; p < p1; ++p) {
    const int32_t fi = indices[p];
    const float x = data[p];
    if (lane == 0) lin = add_rn(lin, mul_rn(ld_gather_f1(W + fi), x));
    const float* vr = V + (size_t)fi * dim;
#pragma unroll
    for (int u = 0; u < UNITS; ++u) {
      const int unit = lane + u * L;
      if (unit < n_units) {
        float vv[VEC];
        if constexpr (VEC == 4) {
          const float4 t4 = ld_gather_f4(reinterpret_cast<const float4*>(vr) + unit);
          vv[0] = t4.x; vv[1] = t4.y; vv[2] = t4.z; vv[3] = t4.w;
        } else {
          vv[0] = ld_gather_f1(vr + unit);
        }
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          const float t = mul_rn(vv[c], x);
          s[u][c] = add_rn(s[u][c], t);
          q[u][c] = add_rn(q[u][c], mul_rn(t, t));
        }
      }
    }
  }
  float inter = 0.0f;
#pragma unroll
  for (int u = 0; u < UNITS; ++u) {
    const int unit = lane + u * L;
    if (unit < n_units) {
#pragma unroll
      for (int c = 0; c < VEC; ++c) inter = add_rn(inter, sub_rn(mul_rn(s[u][c], s[u][c]), q[u][c]));
      if (sums) {
#pragma unroll
        for (int c = 0; c < VEC; ++c) sums[(size_t)r * dim + unit * VEC + c] = s[u][c];
      }
    }
  }
  // group butterfly (sub-warp groups may be partially populated at the grid's tail: use the group mask)
  const unsigned gmask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << ((threadIdx.x & 31) & ~(L - 1)));
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) inter = add_rn(inter, __shfl_xor_sync(gmask, inter, o, L));
  if (lane == 0) yhat[r] = add_rn(add_rn(*w0, lin), mul_rn(0.5f, inter));
}

}  // namespace tfr

using namespace tfr;

extern "C" int tfr_fm_forward(int64_t n_rows, const int64_t* indptr, const int32_t* indices, const float* data,
                              const float* w0, const float* W, const float* V, int32_t dim, float* yhat, float* sums,
                              void* stream) {
  TFR_CHECK_ARG(n_rows >= 0 && dim > 0);
  if (n_rows == 0) return TFR_OK;
  TFR_CHECK_ARG(indptr && indices && data && w0 && W && V && yhat);
  const RowGeom g = row_geom(dim);
  const int units = (dim / g.vec + g.lanes - 1) / g.lanes;
  const int groups_per_cta = 256 / g.lanes;
  const unsigned grid = (unsigned)((n_rows + groups_per_cta - 1) / groups_per_cta);
  cudaStream_t st = (cudaStream_t)stream;
#define TFR_FM_CASE(VV, LL, UU)                                                                               \
  if (g.vec == VV && g.lanes == LL && units == UU) {                                                          \
    fm_forward_kernel<VV, LL, UU><<<grid, 256, 0, st>>>(n_rows, indptr, indices, data, w0, W, V, dim, yhat, sums); \
    TFR_LAUNCH_CHECK();                                                                                        \
    return TFR_OK;                                                                                             \
  }
  TFR_FM_CASE(4, 1, 1) TFR_FM_CASE(4, 2, 1) TFR_FM_CASE(4, 4, 1) TFR_FM_CASE(4, 8, 1) TFR_FM_CASE(4, 16, 1)
  TFR_FM_CASE(4, 32, 1) TFR_FM_CASE(4, 32, 2) TFR_FM_CASE(4, 32, 4)
  TFR_FM_CASE(1, 1, 1) TFR_FM_CASE(1, 2, 1) TFR_FM_CASE(1, 4, 1) TFR_FM_CASE(1, 8, 1) TFR_FM_CASE(1, 16, 1)
  TFR_FM_CASE(1, 32, 1) TFR_FM_CASE(1, 32, 2) TFR_FM_CASE(1, 32, 4)
#undef TFR_FM_CASE
  set_error("unsupported FM dim %d", dim);
  return TFR_ERR_INVALID;
}
