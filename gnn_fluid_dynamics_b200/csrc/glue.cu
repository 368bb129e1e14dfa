// Finite-volume glue around the processor: face-area normalisation, the FVM integrator,
// the divergence / masked-MSE loss terms with their backward passes, and the rollout state advance.  The reference runs
// these as ~20 small tensor kernels per forward plus boolean-mask indexing (host syncs, sort-based index backward);
// here each is ONE kernel, gathers are fixed-degree (3 faces per cell, <= 2 cells per face) and every reduction is
// deterministic (fixed-order fp64 block partials, finalised by the last block to arrive).
//
//   face_area_norm   utils/normalisation.py:325-344 + nn.BatchNorm1d(1)  (models/Fvgn.py:218)
//   fvm_integrate    models/Fvgn.py:221-255  (chain_flux_dot_product: utils/maths.py:12-20)
//   fvm_divergence   utils/fvm.py:26-37
//   masked_mse       utils/loss.py:55-60 (MSE_per_element_torch with an optional row mask)
//   state_advance    rollout.py:340 + models/Fvgn.py:133-148 / Mgn.py:139-151 + normalisation.py:255-278
#include "common.cuh"

namespace gnnfd {

constexpr int GL_THREADS = 256;
constexpr int GL_MAX_BLOCKS = 1184;   // 8 x 148

static int gl_blocks(int64_t n) {
  int64_t b = (n + GL_THREADS - 1) / GL_THREADS;
  int64_t cap = (int64_t)num_sms() * 8;
  if (cap > GL_MAX_BLOCKS) cap = GL_MAX_BLOCKS;      // the reduction workspace holds GL_MAX_BLOCKS partials
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// Deterministic grid reduction of K doubles per thread: warp shuffles in a fixed tree, one partial per block, and the
// LAST block to arrive (ticket) adds the block partials in index order and calls `fin(total)`; it also resets the
// ticket so the workspace can be reused without clearing.  partials: [gridDim.x * K] doubles; ticket: one uint32.
template <int K, typename Fin>
__device__ __forceinline__ void grid_reduce(double (&v)[K], double *partials, unsigned int *ticket, Fin fin) {
  __shared__ double s_part[GL_THREADS / 32][K];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) s_part[warp][k] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double x = 0.0;
      for (int w = 0; w < GL_THREADS / 32; ++w) x += s_part[w][k];
      partials[(size_t)blockIdx.x * K + k] = x;
    }
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double tot[K];
#pragma unroll
    for (int k = 0; k < K; ++k) tot[k] = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b)
#pragma unroll
      for (int k = 0; k < K; ++k) tot[k] += partials[(size_t)b * K + k];
    *ticket = 0u;
    fin(tot);
  }
}

__device__ __forceinline__ float face_raw(const float *area, const float *vol, const int32_t *row, const int32_t *col,
                                          float dt_mean, int64_t f) {
  const float v = (__ldg(vol + __ldg(row + f)) + __ldg(vol + __ldg(col + f))) / 2.0f;
  return __ldg(area + f) * (dt_mean / v);
}
__device__ __forceinline__ float mean_of(const float *dt, int n) {
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += __ldg(dt + i);
  return s / (float)n;
}

// ------------------------------------------------------------------------------------ face-area BatchNorm
// stats[0] = mean, stats[1] = 1 / sqrt(var + eps) used for the normalisation (batch statistics when training)
__global__ void face_area_stats_kernel(const float *area, const float *vol, const int32_t *row, const int32_t *col,
                                       const float *dt, int n_dt, int64_t E, float eps, float momentum, int n_updates,
                                       float *running_mean, float *running_var, long long *nbt, float *stats,
                                       double *partials, unsigned int *ticket) {
  pdl_entry();
  const float dt_mean = mean_of(dt, n_dt);
  double v[2] = {0.0, 0.0};
  for (int64_t f = (int64_t)blockIdx.x * GL_THREADS + threadIdx.x; f < E; f += (int64_t)gridDim.x * GL_THREADS) {
    const double r = (double)face_raw(area, vol, row, col, dt_mean, f);
    v[0] += r;
    v[1] += r * r;
  }
  grid_reduce<2>(v, partials, ticket, [&](const double(&t)[2]) {
    const double mean = t[0] / (double)E;
    double var = t[1] / (double)E - mean * mean;      // biased variance (what BatchNorm normalises with)
    if (var < 0.0) var = 0.0;
    stats[0] = (float)mean;
    stats[1] = (float)(1.0 / sqrt(var + (double)eps));
    const double unbiased = E > 1 ? var * (double)E / (double)(E - 1) : var;
    float rm = *running_mean, rv = *running_var;
    for (int i = 0; i < n_updates; ++i) {             // the reference normalises twice per training step
      rm = (1.0f - momentum) * rm + momentum * (float)mean;
      rv = (1.0f - momentum) * rv + momentum * (float)unbiased;
    }
    *running_mean = rm;
    *running_var = rv;
    if (nbt != nullptr) *nbt += n_updates;
  });
}

__global__ void face_area_apply_kernel(const float *area, const float *vol, const int32_t *row, const int32_t *col,
                                       const float *dt, int n_dt, int64_t E, const float *stats, const float *running_mean,
                                       const float *running_var, float eps, const float *w, const float *b, float *out) {
  pdl_entry();
  const float dt_mean = mean_of(dt, n_dt);
  const float mean = stats != nullptr ? stats[0] : *running_mean;
  const float rstd = stats != nullptr ? stats[1] : rsqrtf(*running_var + eps);
  const float ww = w != nullptr ? *w : 1.f, bb = b != nullptr ? *b : 0.f;
  for (int64_t f = (int64_t)blockIdx.x * GL_THREADS + threadIdx.x; f < E; f += (int64_t)gridDim.x * GL_THREADS)
    out[f] = (face_raw(area, vol, row, col, dt_mean, f) - mean) * rstd * ww + bb;
}

// d weight = sum g * xhat, d bias = sum g   (xhat recovered from the normalised value: xhat = (y - b) / w is not safe
// for w ~ 0, so it is recomputed from the raw value)
__global__ void face_area_bwd_kernel(const float *area, const float *vol, const int32_t *row, const int32_t *col,
                                     const float *dt, int n_dt, int64_t E, const float *stats, const float *running_mean,
                                     const float *running_var, float eps, const float *g, float *dw, float *db,
                                     double *partials, unsigned int *ticket) {
  pdl_entry();
  const float dt_mean = mean_of(dt, n_dt);
  const float mean = stats != nullptr ? stats[0] : *running_mean;
  const float rstd = stats != nullptr ? stats[1] : rsqrtf(*running_var + eps);
  double v[2] = {0.0, 0.0};
  for (int64_t f = (int64_t)blockIdx.x * GL_THREADS + threadIdx.x; f < E; f += (int64_t)gridDim.x * GL_THREADS) {
    const float xh = (face_raw(area, vol, row, col, dt_mean, f) - mean) * rstd;
    const float gg = __ldg(g + f);
    v[0] += (double)gg * (double)xh;
    v[1] += (double)gg;
  }
  grid_reduce<2>(v, partials, ticket, [&](const double(&t)[2]) {
    *dw = (float)t[0];
    *db = (float)t[1];
  });
}

// ------------------------------------------------------------------------------------ FVM integrator
// acc[c] = -(sum_j u_f (u_f . n_cj) a_f) - (sum_j p_f n_cj a_f) / rho + sum_j d_f     (Fvgn.py:221-255)
// eo rows: (u, v, p, d0, d1); optional div[c] = sum_j (u_f . n_cj) a_f   (fvm.py:26-37)
__global__ void fvm_integrate_fwd_kernel(const float *eo, int ld, const float *area, const float *normal,
                                         const int32_t *cf0, const int32_t *cf1, const int32_t *cf2, int64_t N,
                                         float inv_rho, float *acc, float *div) {
  pdl_entry();
  for (int64_t c = (int64_t)blockIdx.x * GL_THREADS + threadIdx.x; c < N; c += (int64_t)gridDim.x * GL_THREADS) {
    float ax = 0.f, ay = 0.f, px = 0.f, py = 0.f, dx = 0.f, dy = 0.f, dv = 0.f;
    const float *nc = normal + c * 6;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int64_t f = __ldg((j == 0 ? cf0 : j == 1 ? cf1 : cf2) + c);
      const float *r = eo + f * ld;
      const float u = __ldg(r), v = __ldg(r + 1), a = __ldg(area + f);
      const float nx = __ldg(nc + 2 * j), ny = __ldg(nc + 2 * j + 1);
      if (acc != nullptr) {
        const float p = __ldg(r + 2);
        ax += ((u * u) * nx + (u * v) * ny) * a;
        ay += ((v * u) * nx + (v * v) * ny) * a;
        px += p * nx * a;
        py += p * ny * a;
        dx += __ldg(r + 3);
        dy += __ldg(r + 4);
      }
      dv += (u * nx + v * ny) * a;
    }
    if (acc != nullptr) {
      acc[2 * c] = 1.0f * (-ax - px * inv_rho) + dx;
      acc[2 * c + 1] = 1.0f * (-ay - py * inv_rho) + dy;
    }
    if (div != nullptr) div[c] = dv;
  }
}

// ------------------------------------------------------------------------------------- face gather
// out[j][c][:] = t[cf_j[c]][0:w]: the three `t[f_graph.face[j]]` gathers every integrator / divergence of the model
// zoo starts with (Fvgn.py:232-246, Flux.py:186-203, VertPot.py:128-146, Conservative.py:..., fvm.py:26-37) as one launch.
__global__ void gather3_fwd_kernel(const float *t, int ld, int w, const int32_t *cf0, const int32_t *cf1,
                                   const int32_t *cf2, int64_t N, float *out) {
  pdl_entry();
  const int64_t total = 3 * N;
  for (int64_t i = (int64_t)blockIdx.x * GL_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * GL_THREADS) {
    const int j = (int)(i / N);
    const int64_t c = i - (int64_t)j * N;
    const int64_t f = __ldg((j == 0 ? cf0 : j == 1 ? cf1 : cf2) + c);
    const float *r = t + f * ld;
    float *o = out + i * w;
    for (int k = 0; k < w; ++k) o[k] = __ldg(r + k);
  }
}
// transpose (autograd of the gather): one thread per face adds the gradients of the slots that gathered it - its (at most
// two) cells row[f], col[f], slot found by comparing the cell's three face ids; fixed order (row cell first, slots
// ascending), no atomics, no sort: d_t[f][0:w] = sum_{(c, j): cf_j[c] = f} g[j][c][0:w]
__global__ void gather3_bwd_kernel(const float *g, int w, const int32_t *cf0, const int32_t *cf1, const int32_t *cf2,
                                   const int32_t *row, const int32_t *col, int64_t N, int64_t E, float *d_t, int ld_d) {
  pdl_entry();
  for (int64_t f = (int64_t)blockIdx.x * GL_THREADS + threadIdx.x; f < E; f += (int64_t)gridDim.x * GL_THREADS) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int32_t c0 = __ldg(row + f), c1 = __ldg(col + f);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int32_t c = s == 0 ? c0 : c1;
      if (s == 1 && c1 == c0) break;                     // boundary face: self-loop, one cell
      if (c < 0) continue;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (__ldg((j == 0 ? cf0 : j == 1 ? cf1 : cf2) + c) == (int32_t)f) {
          const float *r = g + ((int64_t)j * N + c) * w;
          for (int k = 0; k < w; ++k) acc[k] += __ldg(r + k);
        }
      }
    }
    float *o = d_t + f * ld_d;
    for (int k = 0; k < w; ++k) o[k] = acc[k];
  }
}

// ---------------------------------------------------------------------------------- Flux integrator
// FluxA's integrator (Flux.py:166-206) on the signed per-cell face flux (face_flux_to_cell_flux, utils/fvm.py:96-156):
//   s_cj = +1 if c owns face f = cf[j][c], -1 if c is the neighbour of an interior face, else 0;  phi_cj = phi_f s_cj
//   acc[c] = 1 * (-(sum_j u_f phi_cj k_f) - (sum_j p_f n_cj a_f) / rho) + sum_j d_f
// eo rows: (u, v, p, phi, d0, d1); k = normalised mean(dt) / face volume (normalize_vol_dt), a = normalised face area.
// Forward only (evaluation / rollout); every product and sum is separately rounded in the reference's order, so the
// result equals the tensor expression bit for bit.  cell_flux[c, j] = (phi_f * flux_scale + flux_shift) s_cj (optional).
__global__ void flux_integrate_fwd_kernel(const float *eo, int ld, int flux_col, const float *coeff, const float *area,
                                          const float *normal, const int32_t *cf0, const int32_t *cf1, const int32_t *cf2,
                                          const int32_t *row, const int32_t *col, int64_t N, float inv_rho, float *acc,
                                          float *cell_flux, float flux_scale, float flux_shift) {
  pdl_entry();
  for (int64_t c = (int64_t)blockIdx.x * GL_THREADS + threadIdx.x; c < N; c += (int64_t)gridDim.x * GL_THREADS) {
    float ax = 0.f, ay = 0.f, px = 0.f, py = 0.f, dx = 0.f, dy = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int64_t f = __ldg((j == 0 ? cf0 : j == 1 ? cf1 : cf2) + c);
      const float *r = eo + f * ld;
      const int32_t owner = __ldg(row + f), neigh = __ldg(col + f);
      const bool interior = !(owner == neigh || neigh < 0);
      const float sg = (int64_t)owner == c ? 1.f : ((interior && (int64_t)neigh == c) ? -1.f : 0.f);
      const float phi = __ldg(r + flux_col);
      if (cell_flux != nullptr) cell_flux[3 * c + j] = __fmul_rn(__fadd_rn(__fmul_rn(phi, flux_scale), flux_shift), sg);
      if (acc != nullptr) {
        const float cfl = __fmul_rn(phi, sg), k = __ldg(coeff + f), a = __ldg(area + f);
        const float nx = __ldg(normal + c * 6 + 2 * j), ny = __ldg(normal + c * 6 + 2 * j + 1);
        const float tx = __fmul_rn(__fmul_rn(__ldg(r), cfl), k), ty = __fmul_rn(__fmul_rn(__ldg(r + 1), cfl), k);
        const float p = __ldg(r + 2);
        const float qx = __fmul_rn(__fmul_rn(p, nx), a), qy = __fmul_rn(__fmul_rn(p, ny), a);
        ax = __fadd_rn(ax, tx); ay = __fadd_rn(ay, ty);
        px = __fadd_rn(px, qx); py = __fadd_rn(py, qy);
        const float d0 = __ldg(r + 4), d1 = __ldg(r + 5);
        dx = j == 0 ? d0 : __fadd_rn(dx, d0);
        dy = j == 0 ? d1 : __fadd_rn(dy, d1);
      }
    }
    if (acc != nullptr) {
      // (a tensor divided by a Python scalar is multiplied by the scalar's fp32 reciprocal)
      acc[2 * c] = __fadd_rn(__fmul_rn(1.0f, __fsub_rn(-ax, __fmul_rn(px, inv_rho))), dx);
      acc[2 * c + 1] = __fadd_rn(__fmul_rn(1.0f, __fsub_rn(-ay, __fmul_rn(py, inv_rho))), dy);
    }
  }
}

// transpose: one thread per face; its (at most two) cells are c_edge_index[:, f], the slot of f inside a cell is found
// by comparing the cell's three face ids - a fixed-degree gather, no atomics, fixed order (row cell, then col cell).
// g_acc [N,2] (may be NULL), g_div [N] (may be NULL) -> d_eo [E, ld_g] (columns 0..4, accumulated or written), d_area [E]
__global__ void fvm_integrate_bwd_kernel(const float *eo, int ld, const float *area, const float *normal,
                                         const int32_t *cf0, const int32_t *cf1, const int32_t *cf2, const int32_t *row,
                                         const int32_t *col, int64_t E, float inv_rho, const float *g_acc,
                                         const float *g_div, float *d_eo, int ld_g, int n_cols, float *d_area) {
  pdl_entry();
  for (int64_t f = (int64_t)blockIdx.x * GL_THREADS + threadIdx.x; f < E; f += (int64_t)gridDim.x * GL_THREADS) {
    const float *r = eo + f * ld;
    const float u = __ldg(r), v = __ldg(r + 1), a = __ldg(area + f);
    const float p = g_acc != nullptr ? __ldg(r + 2) : 0.f;
    float du = 0.f, dvv = 0.f, dp = 0.f, dd0 = 0.f, dd1 = 0.f, da = 0.f;
    const int32_t c0 = __ldg(row + f), c1 = __ldg(col + f);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int32_t c = s == 0 ? c0 : c1;
      if (s == 1 && c1 == c0) break;                     // boundary face: self-loop, one cell
      int j = -1;
      if (__ldg(cf0 + c) == (int32_t)f) j = 0;
      else if (__ldg(cf1 + c) == (int32_t)f) j = 1;
      else if (__ldg(cf2 + c) == (int32_t)f) j = 2;
      if (j < 0) continue;
      const float nx = __ldg(normal + (int64_t)c * 6 + 2 * j), ny = __ldg(normal + (int64_t)c * 6 + 2 * j + 1);
      if (g_acc != nullptr) {
        const float gx = __ldg(g_acc + 2 * (int64_t)c), gy = __ldg(g_acc + 2 * (int64_t)c + 1);
        du += gx * (-(2.f * u * nx + v * ny) * a) + gy * (-(v * nx) * a);
        dvv += gx * (-(u * ny) * a) + gy * (-(u * nx + 2.f * v * ny) * a);
        dp += -(gx * nx + gy * ny) * a * inv_rho;
        dd0 += gx;
        dd1 += gy;
        da += gx * (-((u * u) * nx + (u * v) * ny) - p * nx * inv_rho) + gy * (-((v * u) * nx + (v * v) * ny) - p * ny * inv_rho);
      }
      if (g_div != nullptr) {
        const float gd = __ldg(g_div + c);
        du += gd * nx * a;
        dvv += gd * ny * a;
        da += gd * (u * nx + v * ny);
      }
    }
    float *o = d_eo + f * ld_g;
    o[0] = du;
    o[1] = dvv;
    if (n_cols > 2) { o[2] = dp; o[3] = dd0; o[4] = dd1; }
    if (d_area != nullptr) d_area[f] = da;
  }
}

// ------------------------------------------------------------------------------------ masked MSE
// out[0] = sum over unmasked rows, all C columns, of (a - b)^2 / (count * C);  out[1] = count * C
__global__ void masked_mse_fwd_kernel(const float *a, int ld_a, const float *b, int ld_b, const uint8_t *mask, int64_t R, int C,
                                      float *out, double *partials, unsigned int *ticket) {
  pdl_entry();
  double v[2] = {0.0, 0.0};
  for (int64_t r = (int64_t)blockIdx.x * GL_THREADS + threadIdx.x; r < R; r += (int64_t)gridDim.x * GL_THREADS) {
    if (mask != nullptr && !mask[r]) continue;
    for (int c = 0; c < C; ++c) {
      const float d = __ldg(a + r * ld_a + c) - __ldg(b + r * ld_b + c);
      v[0] += (double)(d * d);
    }
    v[1] += (double)C;
  }
  grid_reduce<2>(v, partials, ticket, [&](const double(&t)[2]) {
    out[0] = (float)(t[0] / t[1]);      // 0 / 0 = NaN like torch's mean of an empty selection
    out[1] = (float)t[1];
  });
}
// d_a[r, c] = g * 2 (a - b) / count  on unmasked rows, 0 elsewhere  (written to [R, C] with stride ld_d)
__global__ void masked_mse_bwd_kernel(const float *a, int ld_a, const float *b, int ld_b, const uint8_t *mask, int64_t R, int C,
                                      const float *fwd_out, const float *g, float *d_a, int ld_d) {
  pdl_entry();
  const float s = 2.0f * (*g) / fwd_out[1];
  for (int64_t r = (int64_t)blockIdx.x * GL_THREADS + threadIdx.x; r < R; r += (int64_t)gridDim.x * GL_THREADS) {
    const bool on = mask == nullptr || mask[r];
    for (int c = 0; c < C; ++c)
      d_a[r * ld_d + c] = on ? s * (__ldg(a + r * ld_a + c) - __ldg(b + r * ld_b + c)) : 0.f;
  }
}

// ------------------------------------------------------------------------------------ rollout state advance
// One kernel per autoregressive step for the cell state and one for the face state:
//   vel = has_change ? x_raw[:, 0:2] + delta : delta            (rollout.py:336-340)
//   x_raw[:, 0:2] = vel ; x_norm[:, 0:2] = (vel - mean) / std    (update_features + normalizer.input, next step's input)
__global__ void advance_cells_kernel(float *x_raw, int ld_x, const float *delta, int ld_d, int has_change, int64_t N,
                                     float *x_norm, int ld_n, float m0, float s0, float m1, float s1, float *vel_out) {
  pdl_entry();
  for (int64_t c = (int64_t)blockIdx.x * GL_THREADS + threadIdx.x; c < N; c += (int64_t)gridDim.x * GL_THREADS) {
    float u = __ldg(delta + c * ld_d), v = __ldg(delta + c * ld_d + 1);
    if (has_change) { u += x_raw[c * ld_x]; v += x_raw[c * ld_x + 1]; }
    x_raw[c * ld_x] = u;
    x_raw[c * ld_x + 1] = v;
    if (x_norm != nullptr) { x_norm[c * ld_n] = (u - m0) / s0; x_norm[c * ld_n + 1] = (v - m1) / s1; }
    if (vel_out != nullptr) { vel_out[2 * c] = u; vel_out[2 * c + 1] = v; }
  }
}
//   dv = u[row] - u[col]; dv[bc] = bc_value[bc]; f_raw[:, 0:2] = dv; f_norm[:, 0:2] = (dv - mean) / std
__global__ void advance_faces_kernel(const float *x_raw, int ld_x, const int32_t *row, const int32_t *col,
                                     const uint8_t *bc_mask, const float *bc_value, int ld_bc, int64_t E, float *f_raw,
                                     int ld_f, float *f_norm, int ld_n, float m0, float s0, float m1, float s1) {
  pdl_entry();
  for (int64_t f = (int64_t)blockIdx.x * GL_THREADS + threadIdx.x; f < E; f += (int64_t)gridDim.x * GL_THREADS) {
    float du, dv;
    if (bc_mask != nullptr && bc_mask[f]) {
      du = __ldg(bc_value + f * ld_bc);
      dv = __ldg(bc_value + f * ld_bc + 1);
    } else {
      const int64_t r = __ldg(row + f), c = __ldg(col + f);
      du = x_raw[r * ld_x] - x_raw[c * ld_x];
      dv = x_raw[r * ld_x + 1] - x_raw[c * ld_x + 1];
    }
    f_raw[f * ld_f] = du;
    f_raw[f * ld_f + 1] = dv;
    if (f_norm != nullptr) { f_norm[f * ld_n] = (du - m0) / s0; f_norm[f * ld_n + 1] = (dv - m1) / s1; }
  }
}

}  // namespace gnnfd

using namespace gnnfd;

extern "C" size_t gnnfd_glue_workspace_bytes(void) { return (size_t)GL_MAX_BLOCKS * 4 * sizeof(double) + 64; }

static bool split_ws(void *ws, size_t bytes, double *&partials, unsigned int *&ticket) {
  if (ws == nullptr || bytes < gnnfd_glue_workspace_bytes()) return false;
  ticket = (unsigned int *)ws;
  partials = (double *)((uint8_t *)ws + 64);
  return true;
}

extern "C" int gnnfd_face_area_norm(const float *area, const float *volume, const int32_t *row, const int32_t *col,
                                    const float *dt, int32_t n_dt, int64_t n_faces, const float *bn_weight,
                                    const float *bn_bias, float *running_mean, float *running_var,
                                    int64_t *num_batches_tracked, int32_t training, float momentum, float eps,
                                    int32_t n_updates, float *out, float *stats, void *workspace, size_t workspace_bytes,
                                    void *stream) {
  GNNFD_CHECK_ARG(n_faces >= 0 && n_dt >= 1, "bad sizes");
  if (n_faces == 0) return GNNFD_OK;
  GNNFD_CHECK_ARG(area && volume && row && col && dt && out && running_mean && running_var, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = gl_blocks(n_faces);
  if (training) {
    double *partials;
    unsigned int *ticket;
    GNNFD_CHECK_ARG(stats != nullptr, "training mode needs the stats output");
    if (!split_ws(workspace, workspace_bytes, partials, ticket)) { set_error("gnnfd_face_area_norm: workspace too small"); return GNNFD_E_WORKSPACE; }
    launch_pdl(face_area_stats_kernel, dim3(grid), dim3(GL_THREADS), 0, st, area, volume, row, col, dt, n_dt, n_faces, eps, momentum, n_updates,
                                                       running_mean, running_var, (long long *)num_batches_tracked, stats,
                                                       partials, ticket);
    GNNFD_LAUNCH_CHECK();
  }
  launch_pdl(face_area_apply_kernel, dim3(grid), dim3(GL_THREADS), 0, st, area, volume, row, col, dt, n_dt, n_faces, training ? stats : nullptr,
                                                     running_mean, running_var, eps, bn_weight, bn_bias, out);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

extern "C" int gnnfd_face_area_norm_backward(const float *area, const float *volume, const int32_t *row, const int32_t *col,
                                             const float *dt, int32_t n_dt, int64_t n_faces, const float *stats,
                                             const float *running_mean, const float *running_var, float eps, const float *g,
                                             float *d_weight, float *d_bias, void *workspace, size_t workspace_bytes,
                                             void *stream) {
  GNNFD_CHECK_ARG(n_faces > 0 && n_dt >= 1 && area && volume && row && col && dt && g && d_weight && d_bias, "bad arguments");
  GNNFD_CHECK_ARG(stats != nullptr || (running_mean && running_var), "need batch stats or running stats");
  double *partials;
  unsigned int *ticket;
  if (!split_ws(workspace, workspace_bytes, partials, ticket)) { set_error("gnnfd_face_area_norm_backward: workspace too small"); return GNNFD_E_WORKSPACE; }
  launch_pdl(face_area_bwd_kernel, dim3(gl_blocks(n_faces)), dim3(GL_THREADS), 0, (cudaStream_t)stream, 
      area, volume, row, col, dt, n_dt, n_faces, stats, running_mean, running_var, eps, g, d_weight, d_bias, partials, ticket);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

extern "C" int gnnfd_gather3(const float *t, int32_t ld, int32_t width, const int32_t *cf0, const int32_t *cf1,
                             const int32_t *cf2, int64_t n_cells, float *out, void *stream) {
  GNNFD_CHECK_ARG(n_cells >= 0 && width >= 1 && width <= 8 && ld >= width, "bad sizes (width 1..8)");
  if (n_cells == 0) return GNNFD_OK;
  GNNFD_CHECK_ARG(t && cf0 && cf1 && cf2 && out, "null pointer");
  launch_pdl(gather3_fwd_kernel, dim3(gl_blocks(3 * n_cells)), dim3(GL_THREADS), 0, (cudaStream_t)stream, t, ld, width, cf0,
             cf1, cf2, n_cells, out);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

extern "C" int gnnfd_gather3_backward(const float *g, int32_t width, const int32_t *cf0, const int32_t *cf1,
                                      const int32_t *cf2, const int32_t *row, const int32_t *col, int64_t n_cells,
                                      int64_t n_faces, float *d_t, int32_t ld_d, void *stream) {
  GNNFD_CHECK_ARG(n_cells >= 0 && n_faces >= 0 && width >= 1 && width <= 8 && ld_d >= width, "bad sizes (width 1..8)");
  if (n_faces == 0) return GNNFD_OK;
  GNNFD_CHECK_ARG(g && cf0 && cf1 && cf2 && row && col && d_t, "null pointer");
  launch_pdl(gather3_bwd_kernel, dim3(gl_blocks(n_faces)), dim3(GL_THREADS), 0, (cudaStream_t)stream, g, width, cf0, cf1,
             cf2, row, col, n_cells, n_faces, d_t, ld_d);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

extern "C" int gnnfd_flux_integrate(const float *edge_out, int32_t ld, int32_t flux_col, const float *coeff, const float *area,
                                    const float *normal, const int32_t *cf0, const int32_t *cf1, const int32_t *cf2,
                                    const int32_t *row, const int32_t *col, int64_t n_cells, float rho, float *acc,
                                    float *cell_flux, float flux_scale, float flux_shift, void *stream) {
  GNNFD_CHECK_ARG(n_cells >= 0 && ld >= 1 && flux_col >= 0 && flux_col < ld && rho != 0.f, "bad sizes");
  if (n_cells == 0) return GNNFD_OK;
  GNNFD_CHECK_ARG(edge_out && cf0 && cf1 && cf2 && row && col && (acc || cell_flux), "null pointer");
  GNNFD_CHECK_ARG(acc == nullptr || (ld >= 6 && flux_col == 3 && coeff && area && normal),
                  "the integrator reads 6 columns (u, v, p, phi, d0, d1) and needs coeff / area / normal");
  launch_pdl(flux_integrate_fwd_kernel, dim3(gl_blocks(n_cells)), dim3(GL_THREADS), 0, (cudaStream_t)stream, edge_out, ld,
             flux_col, coeff, area, normal, cf0, cf1, cf2, row, col, n_cells, 1.0f / rho, acc, cell_flux, flux_scale, flux_shift);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

extern "C" int gnnfd_fvm_integrate(const float *edge_out, int32_t ld, const float *area, const float *normal,
                                   const int32_t *cf0, const int32_t *cf1, const int32_t *cf2, int64_t n_cells, float rho,
                                   float *acc, float *div, void *stream) {
  GNNFD_CHECK_ARG(n_cells >= 0 && ld >= 2 && rho != 0.f, "bad sizes");
  if (n_cells == 0) return GNNFD_OK;
  GNNFD_CHECK_ARG(edge_out && area && normal && cf0 && cf1 && cf2 && (acc || div), "null pointer");
  GNNFD_CHECK_ARG(acc == nullptr || ld >= 5, "the integrator reads 5 columns (u, v, p, d0, d1)");
  launch_pdl(fvm_integrate_fwd_kernel, dim3(gl_blocks(n_cells)), dim3(GL_THREADS), 0, (cudaStream_t)stream, edge_out, ld, area, normal, cf0, cf1,
                                                                                       cf2, n_cells, 1.0f / rho, acc, div);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

extern "C" int gnnfd_fvm_integrate_backward(const float *edge_out, int32_t ld, const float *area, const float *normal,
                                            const int32_t *cf0, const int32_t *cf1, const int32_t *cf2, const int32_t *row,
                                            const int32_t *col, int64_t n_faces, float rho, const float *g_acc,
                                            const float *g_div, float *d_edge_out, int32_t ld_g, int32_t n_cols,
                                            float *d_area, void *stream) {
  GNNFD_CHECK_ARG(n_faces >= 0 && ld >= 2 && rho != 0.f && (n_cols == 2 || n_cols == 5) && ld_g >= n_cols, "bad sizes");
  if (n_faces == 0) return GNNFD_OK;
  GNNFD_CHECK_ARG(edge_out && area && normal && cf0 && cf1 && cf2 && row && col && d_edge_out && (g_acc || g_div), "null pointer");
  GNNFD_CHECK_ARG(g_acc == nullptr || (ld >= 5 && n_cols == 5), "the integrator's backward writes 5 columns");
  launch_pdl(fvm_integrate_bwd_kernel, dim3(gl_blocks(n_faces)), dim3(GL_THREADS), 0, (cudaStream_t)stream, 
      edge_out, ld, area, normal, cf0, cf1, cf2, row, col, n_faces, 1.0f / rho, g_acc, g_div, d_edge_out, ld_g, n_cols, d_area);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

extern "C" int gnnfd_masked_mse(const float *a, int32_t ld_a, const float *b, int32_t ld_b, const uint8_t *mask, int64_t rows,
                                int32_t cols, float *out2, void *workspace, size_t workspace_bytes, void *stream) {
  GNNFD_CHECK_ARG(rows >= 0 && cols >= 1 && ld_a >= cols && ld_b >= cols && a && b && out2, "bad arguments");
  double *partials;
  unsigned int *ticket;
  if (!split_ws(workspace, workspace_bytes, partials, ticket)) { set_error("gnnfd_masked_mse: workspace too small"); return GNNFD_E_WORKSPACE; }
  launch_pdl(masked_mse_fwd_kernel, dim3(gl_blocks(rows)), dim3(GL_THREADS), 0, (cudaStream_t)stream, a, ld_a, b, ld_b, mask, rows, cols, out2,
                                                                                 partials, ticket);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

extern "C" int gnnfd_masked_mse_backward(const float *a, int32_t ld_a, const float *b, int32_t ld_b, const uint8_t *mask,
                                         int64_t rows, int32_t cols, const float *fwd_out2, const float *g, float *d_a,
                                         int32_t ld_d, void *stream) {
  GNNFD_CHECK_ARG(rows >= 0 && cols >= 1 && ld_a >= cols && ld_b >= cols && ld_d >= cols && a && b && fwd_out2 && g && d_a,
                  "bad arguments");
  if (rows == 0) return GNNFD_OK;
  launch_pdl(masked_mse_bwd_kernel, dim3(gl_blocks(rows)), dim3(GL_THREADS), 0, (cudaStream_t)stream, a, ld_a, b, ld_b, mask, rows, cols, fwd_out2,
                                                                                 g, d_a, ld_d);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

// t[r, cols[k]] = (t[r, cols[k]] - shift[k]) / scale[k]   (inverse: t * scale + shift; two roundings each, like the
// tensor expressions they replace)
__global__ void __launch_bounds__(GL_THREADS) affine_columns_kernel(float *__restrict__ t, int64_t rows, int ld, int n_spec,
                                                                  const int32_t *__restrict__ cols,
                                                                  const float *__restrict__ shift,
                                                                  const float *__restrict__ scale, int inverse) {
  pdl_entry();
  const int64_t total = rows * n_spec;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / n_spec;
    const int k = (int)(i - r * n_spec);
    float *p = t + r * ld + __ldg(cols + k);
    const float a = __ldg(shift + k), b = __ldg(scale + k), x = *p;
    *p = inverse ? __fadd_rn(__fmul_rn(x, b), a) : __fdiv_rn(__fsub_rn(x, a), b);
  }
}

extern "C" int gnnfd_affine_columns(float *t, int64_t rows, int32_t ld, int32_t n_spec, const int32_t *cols,
                                    const float *shift, const float *scale, int32_t inverse, void *stream) {
  GNNFD_CHECK_ARG(rows >= 0 && n_spec >= 0 && ld >= 1, "bad sizes");
  if (rows == 0 || n_spec == 0) return GNNFD_OK;
  GNNFD_CHECK_ARG(t && cols && shift && scale, "null pointer");
  launch_pdl(affine_columns_kernel, dim3(gl_blocks(rows * n_spec)), dim3(GL_THREADS), 0, (cudaStream_t)stream, t, rows, ld, n_spec, cols, shift,
                                                                                          scale, inverse);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

extern "C" int gnnfd_state_advance(float *x_raw, int32_t ld_x, const float *delta, int32_t ld_d, int32_t has_change,
                                   int64_t n_cells, float *x_norm, int32_t ld_xn, const float *cell_mean_std4,
                                   const int32_t *row, const int32_t *col, const uint8_t *bc_mask, const float *bc_value,
                                   int32_t ld_bc, int64_t n_faces, float *f_raw, int32_t ld_f, float *f_norm, int32_t ld_fn,
                                   const float *face_mean_std4, float *vel_out, void *stream) {
  GNNFD_CHECK_ARG(n_cells >= 0 && n_faces >= 0 && x_raw && delta && ld_x >= 2 && ld_d >= 2, "bad arguments");
  GNNFD_CHECK_ARG(x_norm == nullptr || cell_mean_std4 != nullptr, "x_norm needs (mean0, std0, mean1, std1)");
  GNNFD_CHECK_ARG(f_norm == nullptr || face_mean_std4 != nullptr, "f_norm needs (mean0, std0, mean1, std1)");
  cudaStream_t st = (cudaStream_t)stream;
  const float *cm = cell_mean_std4, *fm = face_mean_std4;
  if (n_cells > 0) {
    launch_pdl(advance_cells_kernel, dim3(gl_blocks(n_cells)), dim3(GL_THREADS), 0, st, x_raw, ld_x, delta, ld_d, has_change, n_cells, x_norm, ld_xn,
                                                                   cm ? cm[0] : 0.f, cm ? cm[1] : 1.f, cm ? cm[2] : 0.f,
                                                                   cm ? cm[3] : 1.f, vel_out);
    GNNFD_LAUNCH_CHECK();
  }
  if (n_faces > 0 && f_raw != nullptr) {
    GNNFD_CHECK_ARG(row && col && ld_f >= 2, "face update needs row / col");
    GNNFD_CHECK_ARG(bc_mask == nullptr || (bc_value != nullptr && ld_bc >= 2), "bc_mask needs bc_value");
    launch_pdl(advance_faces_kernel, dim3(gl_blocks(n_faces)), dim3(GL_THREADS), 0, st, x_raw, ld_x, row, col, bc_mask, bc_value, ld_bc, n_faces,
                                                                   f_raw, ld_f, f_norm, ld_fn, fm ? fm[0] : 0.f,
                                                                   fm ? fm[1] : 1.f, fm ? fm[2] : 0.f, fm ? fm[3] : 1.f);
    GNNFD_LAUNCH_CHECK();
  }
  return GNNFD_OK;
}
