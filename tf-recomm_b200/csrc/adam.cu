// K4: TF AdamOptimizer's IndexedSlices semantics (TF: adam.py::_apply_sparse_shared; the
// optimizer north_star names, selected at ops.py:144/148):
//     m   = m*beta1                (WHOLE table)      m[uq]  += g*(1-beta1)
//     v   = v*beta2                (WHOLE table)      v[uq]  += (g*g)*(1-beta2)
//     var = var - (lr_t*m)/(sqrt(v)+eps)  (WHOLE table)
// TF runs this as ~8 unfused table-wide kernels; here each parameter is read and written exactly once
// per step (24 B/param): adam_stream_kernel is the streaming pass over the rows NOT in this step's
// slice (g = 0), adam_touched_kernel updates the slice rows with their summed gradient.  Every
// operation keeps TF's rounding order (separate mul/add, correctly rounded sqrt and divide).
// Also: dense ApplyAdam for bias_global (TF: training_ops.cc, A.5), scatter_sub SGD (ops.py:145) and
// the end-of-step bookkeeping (TF: adam.py::_finish).
#include "common.cuh"

namespace tfr {

struct AdamK {
  float b1, b2, lr_t, eps, omb1, omb2;
};
__device__ __forceinline__ AdamK load_k(const tfr_opt_scalars* opt) {
  AdamK k;
  k.b1 = opt->beta1; k.b2 = opt->beta2; k.lr_t = opt->lr_t; k.eps = opt->eps;
  k.omb1 = opt->one_minus_beta1; k.omb2 = opt->one_minus_beta2;
  return k;
}
__device__ __forceinline__ void adam_decay(float& var, float& m, float& v, const AdamK& k) {
  m = mul_rn(m, k.b1);
  v = mul_rn(v, k.b2);
  var = sub_rn(var, div_rn(mul_rn(k.lr_t, m), add_rn(sqrt_rn(v), k.eps)));
}
__device__ __forceinline__ void adam_grad(float& var, float& m, float& v, float g, const AdamK& k) {
  m = add_rn(mul_rn(m, k.b1), mul_rn(g, k.omb1));
  v = add_rn(mul_rn(v, k.b2), mul_rn(mul_rn(g, g), k.omb2));
  var = sub_rn(var, div_rn(mul_rn(k.lr_t, m), add_rn(sqrt_rn(v), k.eps)));
}

// ---- streaming pass, 128-bit path: width % 4 == 0, so a float4 never straddles two rows --------------
// Persistent grid (a multiple of the SM count); each thread owns UNROLL float4 triples per trip so
// that 6*UNROLL 16-byte requests are in flight per thread.  Loads bypass L1 and are evict-first in L2:
// every byte is touched once per step and must not push the batch's gathered rows out of L2.
template <typename IdxT, int UNROLL>
__global__ void __launch_bounds__(512) adam_stream_vec4_kernel(float4* __restrict__ var, float4* __restrict__ m,
                                                               float4* __restrict__ v, IdxT n4, uint32_t row4,
                                                               const uint8_t* __restrict__ touched,
                                                               const tfr_opt_scalars* __restrict__ opt) {
  const AdamK k = load_k(opt);
  const IdxT stride = (IdxT)gridDim.x * blockDim.x;
  for (IdxT q0 = (IdxT)blockIdx.x * blockDim.x + threadIdx.x; q0 < n4; q0 += stride * UNROLL) {
    bool live[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const IdxT q = q0 + (IdxT)u * stride;
      live[u] = q < n4 && touched[q / row4] == 0;
    }
    float4 a[UNROLL], b[UNROLL], c[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const IdxT q = q0 + (IdxT)u * stride;
      if (live[u]) { a[u] = ld_stream_f4(var + q); b[u] = ld_stream_f4(m + q); c[u] = ld_stream_f4(v + q); }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (!live[u]) continue;
      const IdxT q = q0 + (IdxT)u * stride;
      adam_decay(a[u].x, b[u].x, c[u].x, k);
      adam_decay(a[u].y, b[u].y, c[u].y, k);
      adam_decay(a[u].z, b[u].z, c[u].z, k);
      adam_decay(a[u].w, b[u].w, c[u].w, k);
      st_stream_f4(var + q, a[u]);
      st_stream_f4(m + q, b[u]);
      st_stream_f4(v + q, c[u]);
    }
  }
}

// scalar path: any width (dim = 15 rows are 60 B; bias tables have width 1)
template <typename IdxT>
__global__ void __launch_bounds__(512) adam_stream_scalar_kernel(float* __restrict__ var, float* __restrict__ m,
                                                                 float* __restrict__ v, IdxT n, uint32_t width,
                                                                 const uint8_t* __restrict__ touched,
                                                                 const tfr_opt_scalars* __restrict__ opt) {
  const AdamK k = load_k(opt);
  const IdxT stride = (IdxT)gridDim.x * blockDim.x;
  for (IdxT j = (IdxT)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    if (touched[j / width]) continue;
    float a = ld_stream_f1(var + j), b = ld_stream_f1(m + j), c = ld_stream_f1(v + j);
    adam_decay(a, b, c, k);
    st_stream_f1(var + j, a);
    st_stream_f1(m + j, b);
    st_stream_f1(v + j, c);
  }
}

// ---- slice rows: one lane group per sorted entry; only run heads (first entry of a run) act --------
template <int VEC, int L>
__global__ void __launch_bounds__(256) adam_touched_kernel(float* __restrict__ var, float* __restrict__ m,
                                                           float* __restrict__ v, int width,
                                                           const int32_t* __restrict__ sid, int64_t n,
                                                           const float* __restrict__ gsum,
                                                           const tfr_opt_scalars* __restrict__ opt, int sgd) {
  const int lane = threadIdx.x & (L - 1);
  const int64_t kk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  if (kk >= n) return;
  const int32_t id = sid[kk];
  if (kk > 0 && sid[kk - 1] == id) return;
  const size_t base = (size_t)id * width;
  const float* g = gsum + (size_t)kk * width;
  if (sgd) {  // ops.py:145: scatter_sub; gsum already holds the in-order sum of lr*g
    for (int c = lane * VEC; c < width; c += L * VEC) {
      if constexpr (VEC == 4) {
        float4 a = *reinterpret_cast<float4*>(var + base + c);
        const float4 gg = *reinterpret_cast<const float4*>(g + c);
        a.x = sub_rn(a.x, gg.x); a.y = sub_rn(a.y, gg.y); a.z = sub_rn(a.z, gg.z); a.w = sub_rn(a.w, gg.w);
        *reinterpret_cast<float4*>(var + base + c) = a;
      } else {
        var[base + c] = sub_rn(var[base + c], g[c]);
      }
    }
    return;
  }
  const AdamK k = load_k(opt);
  for (int c = lane * VEC; c < width; c += L * VEC) {
    if constexpr (VEC == 4) {
      float4 a = *reinterpret_cast<float4*>(var + base + c);
      float4 b = *reinterpret_cast<float4*>(m + base + c);
      float4 d = *reinterpret_cast<float4*>(v + base + c);
      const float4 gg = *reinterpret_cast<const float4*>(g + c);
      adam_grad(a.x, b.x, d.x, gg.x, k);
      adam_grad(a.y, b.y, d.y, gg.y, k);
      adam_grad(a.z, b.z, d.z, gg.z, k);
      adam_grad(a.w, b.w, d.w, gg.w, k);
      *reinterpret_cast<float4*>(var + base + c) = a;
      *reinterpret_cast<float4*>(m + base + c) = b;
      *reinterpret_cast<float4*>(v + base + c) = d;
    } else {
      float a = var[base + c], b = m[base + c], d = v[base + c];
      adam_grad(a, b, d, g[c], k);
      var[base + c] = a; m[base + c] = b; v[base + c] = d;
    }
  }
}

// ---- end of step ----------------------------------------------------------------------------------
// every CTA clears the touched marks of its slice of the batch; CTA 0 / warp 0 folds the per-CTA
// partials in a fixed order, applies the dense update of bias_global and advances the step scalars.
__global__ void __launch_bounds__(256) finish_step_kernel(tfr_svd_tables t, tfr_opt_scalars* opt,
                                                          const int32_t* __restrict__ users,
                                                          const int32_t* __restrict__ items, int64_t B,
                                                          const float* __restrict__ partials,
                                                          const double* __restrict__ se_partials, int n_partials) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    t.user_touched[users[b]] = 0;
    t.item_touched[items[b]] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    float a = 0.0f;
    double se = 0.0;
    for (int j = threadIdx.x; j < n_partials; j += 32) { a = add_rn(a, partials[j]); se += se_partials[j]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a = add_rn(a, __shfl_xor_sync(0xffffffffu, a, o));
      se += __shfl_xor_sync(0xffffffffu, se, o);
    }
    if (threadIdx.x == 0) {
      const float g = a;  // d cost / d bias_global = sum_b e_b  (A.3)
      opt->g_mu = g;
      opt->se_sum = se;
      if (opt->se_ring && opt->se_ring_len > 0) opt->se_ring[opt->global_step % opt->se_ring_len] = se;
      const bool sgd = opt->flags & TFR_OPT_SGD;
      if (opt->var_mask & TFR_VAR_MU) {
        if (sgd) {
          *t.mu = sub_rn(*t.mu, mul_rn(opt->lr, g));
        } else {  // TF: training_ops.cc ApplyAdam (A.5)
          float alpha = sqrt_rn(sub_rn(1.0f, opt->beta2_power));
          alpha = mul_rn(opt->lr, alpha);
          alpha = div_rn(alpha, sub_rn(1.0f, opt->beta1_power));
          float mm = *t.m_mu, vv = *t.v_mu;
          mm = add_rn(mm, mul_rn(sub_rn(g, mm), opt->one_minus_beta1));
          vv = add_rn(vv, mul_rn(sub_rn(mul_rn(g, g), vv), opt->one_minus_beta2));
          *t.m_mu = mm;
          *t.v_mu = vv;
          *t.mu = sub_rn(*t.mu, div_rn(mul_rn(mm, alpha), add_rn(sqrt_rn(vv), opt->eps)));
        }
      }
      if (!sgd) {  // TF: adam.py::_finish
        opt->beta1_power = mul_rn(opt->beta1_power, opt->beta1);
        opt->beta2_power = mul_rn(opt->beta2_power, opt->beta2);
      }
      opt->global_step += 1;
      opt->batch_cursor += 1;
    }
  }
}

__global__ void opt_init_kernel(tfr_opt_scalars* opt, float lr, float reg, float beta1, float beta2, float eps,
                                int flags, int var_mask) {
  opt->lr = lr; opt->reg = reg; opt->beta1 = beta1; opt->beta2 = beta2; opt->eps = eps;
  opt->beta1_power = beta1; opt->beta2_power = beta2;  // TF: adam.py _create_slots
  opt->one_minus_beta1 = sub_rn(1.0f, beta1);
  opt->one_minus_beta2 = sub_rn(1.0f, beta2);
  float tt = sqrt_rn(sub_rn(1.0f, beta2));
  opt->lr_t = div_rn(mul_rn(lr, tt), sub_rn(1.0f, beta1));
  opt->flags = flags; opt->var_mask = var_mask;
  opt->global_step = 0; opt->batch_cursor = 0;
  opt->se_sum = 0.0; opt->g_mu = 0.0f; opt->pad_ = 0.0f;
  opt->se_ring = nullptr; opt->se_ring_len = 0;
}

}  // namespace tfr

using namespace tfr;

extern "C" int tfr_opt_init(tfr_opt_scalars* opt_dev, float lr, float reg, float beta1, float beta2, float eps,
                            int32_t flags, int32_t var_mask, void* stream) {
  TFR_CHECK_ARG(opt_dev);
  opt_init_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(opt_dev, lr, reg, beta1, beta2, eps, flags, var_mask);
  TFR_LAUNCH_CHECK();
  return TFR_OK;
}

extern "C" int tfr_adam_stream_untouched(float* var, float* m, float* v, int64_t rows, int32_t width,
                                         const uint8_t* touched, const tfr_opt_scalars* opt, void* stream) {
  TFR_CHECK_ARG(rows >= 0 && width > 0);
  if (rows == 0) return TFR_OK;
  TFR_CHECK_ARG(var && m && v && touched && opt);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = rows * (int64_t)width;
  const int sms = sm_count();
  if (width % 4 == 0 && ((uintptr_t)var % 16 == 0) && ((uintptr_t)m % 16 == 0) && ((uintptr_t)v % 16 == 0)) {
    const int64_t n4 = n / 4;
    constexpr int UNROLL = 2;
    int64_t grid = (n4 + 512 * UNROLL - 1) / (512 * UNROLL);
    const int64_t cap = (int64_t)sms * 3;  // 3 CTAs x 512 threads per SM
    if (grid > cap) grid = cap;
    if (n4 < ((int64_t)1 << 31))
      adam_stream_vec4_kernel<uint32_t, UNROLL><<<(unsigned)grid, 512, 0, st>>>(
          (float4*)var, (float4*)m, (float4*)v, (uint32_t)n4, (uint32_t)(width / 4), touched, opt);
    else
      adam_stream_vec4_kernel<uint64_t, UNROLL><<<(unsigned)grid, 512, 0, st>>>(
          (float4*)var, (float4*)m, (float4*)v, (uint64_t)n4, (uint32_t)(width / 4), touched, opt);
  } else {
    int64_t grid = (n + 511) / 512;
    const int64_t cap = (int64_t)sms * 4;
    if (grid > cap) grid = cap;
    if (n < ((int64_t)1 << 31))
      adam_stream_scalar_kernel<uint32_t><<<(unsigned)grid, 512, 0, st>>>(var, m, v, (uint32_t)n, (uint32_t)width,
                                                                        touched, opt);
    else
      adam_stream_scalar_kernel<uint64_t><<<(unsigned)grid, 512, 0, st>>>(var, m, v, (uint64_t)n, (uint32_t)width,
                                                                        touched, opt);
  }
  TFR_LAUNCH_CHECK();
  return TFR_OK;
}

static int launch_touched(float* var, float* m, float* v, int32_t width, const int32_t* sorted_ids, int64_t n,
                          const float* gsum, const tfr_opt_scalars* opt, int sgd, cudaStream_t st) {
  const RowGeom g = row_geom(width);
  const int groups_per_cta = 256 / g.lanes;
  const unsigned grid = (unsigned)((n + groups_per_cta - 1) / groups_per_cta);
#define TFR_TOUCH_CASE(V, LL)                                                                                 \
  if (g.vec == V && g.lanes == LL) {                                                                          \
    adam_touched_kernel<V, LL><<<grid, 256, 0, st>>>(var, m, v, width, sorted_ids, n, gsum, opt, sgd);        \
    TFR_LAUNCH_CHECK();                                                                                        \
    return TFR_OK;                                                                                             \
  }
  TFR_TOUCH_CASE(4, 1) TFR_TOUCH_CASE(4, 2) TFR_TOUCH_CASE(4, 4) TFR_TOUCH_CASE(4, 8) TFR_TOUCH_CASE(4, 16)
  TFR_TOUCH_CASE(4, 32) TFR_TOUCH_CASE(1, 1) TFR_TOUCH_CASE(1, 2) TFR_TOUCH_CASE(1, 4) TFR_TOUCH_CASE(1, 8)
  TFR_TOUCH_CASE(1, 16) TFR_TOUCH_CASE(1, 32)
#undef TFR_TOUCH_CASE
  set_error("unsupported width %d", width);
  return TFR_ERR_INVALID;
}

extern "C" int tfr_adam_touched(float* var, float* m, float* v, int32_t width, const int32_t* sorted_ids, int64_t n,
                                const float* gsum, const tfr_opt_scalars* opt, void* stream) {
  TFR_CHECK_ARG(n >= 0 && width > 0);
  if (n == 0) return TFR_OK;
  TFR_CHECK_ARG(var && m && v && sorted_ids && gsum && opt);
  return launch_touched(var, m, v, width, sorted_ids, n, gsum, opt, 0, (cudaStream_t)stream);
}

extern "C" int tfr_sgd_apply(float* var, int32_t width, const int32_t* sorted_ids, int64_t n, const float* gsum,
                             void* stream) {
  TFR_CHECK_ARG(n >= 0 && width > 0);
  if (n == 0) return TFR_OK;
  TFR_CHECK_ARG(var && sorted_ids && gsum);
  return launch_touched(var, nullptr, nullptr, width, sorted_ids, n, gsum, nullptr, 1, (cudaStream_t)stream);
}

extern "C" int tfr_svd_finish_step(const tfr_svd_tables* t, tfr_opt_scalars* opt, const int32_t* users,
                                   const int32_t* items, int64_t B, const tfr_svd_step_ws* ws, int32_t n_partials,
                                   void* stream) {
  TFR_CHECK_ARG(t && opt && users && items && ws && B > 0 && n_partials > 0 && n_partials <= TFR_MAX_PARTIALS);
  finish_step_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*t, opt, users, items, B,
                                                                                  ws->partials, ws->se_partials,
                                                                                  n_partials);
  TFR_LAUNCH_CHECK();
  return TFR_OK;
}
