/*
 * CPU restatement in C of the DuaLip matching objective's calculate().  TEST INFRASTRUCTURE ONLY: used by tests/
 * as a fast checker at sizes the numpy oracle cannot reach, and by bench.py as the CPU baseline ("port").
 * The CUDA product never links or calls this file.
 *
 * Follows reference src/dualip/objectives/matching.py:116-188 per column at its true length, with the zero padding of
 * the reference's [L x K] blocks restated literally where it changes the result (simplex_eq: the sorted scan runs
 * over L positions, the last L - d of them zeros; L per (class, length bucket) is passed in by the caller):
 *   v = fl(fl(a * fl(s*lambda_r)) + fl(s*c)),  s = fl32(-1/gamma)          matching.py:133-142
 *   box/cone: x = min(max(v,lo),hi)                                          projections/box.py:16, cone.py:21-28
 *   simplex : u = max(v,0); feasible / top-2 shortcut / sorted scan          projections/simplex.py:143-236
 *   grad_r += fl(a*x) ; cx += c*x ; xx += x*x                               matching.py:153-160
 * For "simplex" the padded zeros of the reference's [L x K] blocks never change the result except that a 1-entry
 * column in a bucket of padded length 1 skips the shortcut (simplex.py:166); flag bit 0 of the class says so.
 * Prefix sums over the sorted column are accumulated in double and rounded per element, like torch's CPU cumsum.
 * Compile WITHOUT -ffast-math and with -ffp-contract=off (see oracle/Makefile).
 *
 * Pinned by tests/test_oracle_golden.py against tests/golden/ (outputs of the reference itself).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int32_t kind; /* 0 clamp, 1 simplex (ineq), 2 simplex_eq */
  float lo, hi, z, z_thr;
  uint32_t flags; /* bit0: 1-entry columns are unpadded (L == 1) */
} oracle_class;

static int cmp_desc(const void* pa, const void* pb) {
  const float a = *(const float*)pa, b = *(const float*)pb;
  return (a < b) - (a > b);
}

/* Projects one column in place: v[0..d) -> x[0..d).  Returns branch (0/1/2) and *rho. scratch holds d floats.
 * d_pad >= d: length of the zero-padded column the reference's scan sees (sparse_utils.py:207-208). */
static int project_column(float* v, int64_t d, int64_t d_pad, const oracle_class* pc, float* scratch, int* rho) {
  *rho = 0;
  if (pc->kind == 0) {
    for (int64_t k = 0; k < d; ++k) v[k] = fminf(fmaxf(v[k], pc->lo), pc->hi);
    return -1;
  }
  const float z = pc->z;
  float sum = 0.0f, m1 = 0.0f, m2 = 0.0f;
  int64_t am = 0;
  int have = 0;
  for (int64_t k = 0; k < d; ++k) {
    const float u = fmaxf(v[k], 0.0f);
    v[k] = u;
    sum = sum + u;
    const float un = u / z;
    if (!have || un > m1) {
      if (have) m2 = m1;
      m1 = un;
      am = k;
      have = 1;
    } else if (un > m2) {
      m2 = un;
    }
  }
  /* zero padding takes part in the top-2 of the reference's block: second value is at least 0 */
  if (pc->kind == 1 && sum <= pc->z_thr) return 0;
  const int padded = (d > 1) || !(pc->flags & 1u);
  if (padded && (m1 - m2) > 1.0f) {
    for (int64_t k = 0; k < d; ++k) v[k] = (k == am) ? z : 0.0f;
    *rho = 1;
    return 1;
  }
  memcpy(scratch, v, (size_t)d * sizeof(float));
  qsort(scratch, (size_t)d, sizeof(float), cmp_desc);
  double acc = 0.0;
  int64_t r = 0;
  float css_r = 0.0f;
  for (int64_t i = 0; i < d_pad; ++i) {
    const float ui = i < d ? scratch[i] : 0.0f; /* padded zeros sort to the end */
    acc += (double)ui;
    const float css = (float)acc;
    const float t = (css - z) / (float)(i + 1);
    if (ui - t > 0.0f) {
      r = i + 1;
      css_r = css;
    }
    if (i == 0 && r == 0) css_r = css; /* rho0 = 0 fallback (simplex.py:225) */
  }
  if (r == 0) r = 1;
  const float theta = (css_r - z) / (float)r;
  for (int64_t k = 0; k < d; ++k) v[k] = fmaxf(v[k] - theta, 0.0f);
  *rho = (int)r;
  return 2;
}

/* One evaluation.  grad_out[m] = sum_j a_rj x_rj - b (b may be NULL); scal_out = {dual_obj, cx, reg, lam.grad,
 * max_pos_slack, sum_pos_slack, xx}; x_out (nnz) and diag_out (n_cols, branch | rho<<2, 255 = n/a) may be NULL.
 * col_class may be NULL (all columns class 0).  Returns 0. */
int oracle_matching_calculate(int64_t n_cols, int64_t nnz, int32_t m, const int64_t* ccol, const int64_t* row,
                              const float* a, const float* c, const uint8_t* col_class, const oracle_class* classes,
                              const float* lambda, const float* b, double gamma, float* grad_out, double* scal_out,
                              float* x_out, uint8_t* diag_out, int n_threads, const int32_t* pad_len) {
  /* pad_len: n_classes x 32 padded block lengths indexed by ceil(log2(d)), or NULL (no padding) */
  (void)nnz;
  const float s = (float)(-1.0 / gamma);
  float* sl = (float*)malloc(sizeof(float) * (size_t)m);
  for (int32_t r = 0; r < m; ++r) sl[r] = s * lambda[r];
  int nt = 1;
#ifdef _OPENMP
  nt = n_threads > 0 ? n_threads : omp_get_max_threads();
#endif
  double* gacc = (double*)calloc((size_t)nt * (size_t)m, sizeof(double));
  double cx_tot = 0.0, xx_tot = 0.0;
#pragma omp parallel num_threads(nt) reduction(+ : cx_tot, xx_tot)
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    double* g = gacc + (size_t)tid * (size_t)m;
    int64_t cap = 64;
    float* v = (float*)malloc(sizeof(float) * (size_t)cap);
    float* scratch = (float*)malloc(sizeof(float) * (size_t)cap);
#pragma omp for schedule(static)
    for (int64_t j = 0; j < n_cols; ++j) {
      const int64_t e0 = ccol[j], d = ccol[j + 1] - e0;
      if (diag_out) diag_out[j] = 255;
      if (d <= 0) continue;
      if (d > cap) {
        cap = d * 2;
        v = (float*)realloc(v, sizeof(float) * (size_t)cap);
        scratch = (float*)realloc(scratch, sizeof(float) * (size_t)cap);
      }
      for (int64_t k = 0; k < d; ++k) {
        const float t = a[e0 + k] * sl[row[e0 + k]];
        v[k] = t + s * c[e0 + k];
      }
      const int cls = col_class ? col_class[j] : 0;
      const oracle_class* pc = &classes[cls];
      int64_t d_pad = d;
      if (pad_len && pc->kind == 2) {
        int bkt = 0;
        while (((int64_t)1 << bkt) < d) ++bkt;
        if (pad_len[cls * 32 + bkt] > d_pad) d_pad = pad_len[cls * 32 + bkt];
      }
      int rho = 0;
      const int br = project_column(v, d, d_pad, pc, scratch, &rho);
      if (diag_out && br >= 0) diag_out[j] = (uint8_t)(br | ((rho > 63 ? 63 : rho) << 2));
      for (int64_t k = 0; k < d; ++k) {
        const float x = v[k];
        const float p = a[e0 + k] * x;
        g[row[e0 + k]] += (double)p;
        cx_tot += (double)c[e0 + k] * (double)x;
        xx_tot += (double)x * (double)x;
        if (x_out) x_out[e0 + k] = x;
      }
    }
    free(v);
    free(scratch);
  }
  double lg = 0.0, sp = 0.0;
  float mx = -INFINITY;
  for (int32_t r = 0; r < m; ++r) {
    double t = 0.0;
    for (int k = 0; k < nt; ++k) t += gacc[(size_t)k * (size_t)m + r];
    float gr = (float)t;
    if (b) gr = gr - b[r];
    grad_out[r] = gr;
    lg += (double)lambda[r] * (double)gr;
    if (gr > 0.0f) sp += (double)gr;
    if (gr > mx) mx = gr;
  }
  const double reg = 0.5 * gamma * xx_tot;
  if (scal_out) {
    scal_out[0] = cx_tot + reg + lg;
    scal_out[1] = cx_tot;
    scal_out[2] = reg;
    scal_out[3] = lg;
    scal_out[4] = mx > 0.0f ? (double)mx : 0.0;
    scal_out[5] = sp;
    scal_out[6] = xx_tot;
  }
  free(gacc);
  free(sl);
  return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
