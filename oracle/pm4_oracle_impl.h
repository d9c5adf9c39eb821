/* pm4_oracle_impl.h -- body of the CPU oracle, compiled twice (REAL = float, double).
 *
 * TEST INFRASTRUCTURE ONLY.  See pm4_oracle.c for the header comment and provenance.
 * Every function is suffixed with SFX (f32 / f64) through the NAME() macro.
 */

/* ----------------------------------------------------------------------------------------
 * Log-densities: the TFP log_prob closed forms the reference delegates to
 * (/root/reference/pymc4/distributions/continuous.py:108-112 Normal -> tfd.Normal,
 *  :234-238 HalfNormal, :659-663 HalfCauchy(loc=0); discrete.py:75-78 Bernoulli(probs=),
 *  :500-503 Poisson(rate=)); SURVEY App. B.  Each returns lp and the partial derivatives
 *  w.r.t. value and parameters.
 * -------------------------------------------------------------------------------------- */
static inline REAL NAME(lpdf)(int kind, REAL x, const REAL* q, REAL* dx, REAL* dq) {
  const REAL HALF_LOG_2PI = (REAL)0.91893853320467274178;
  const REAL HALF_LOG_HALF_PI = (REAL)0.22579135264472743236;
  const REAL LOG_2_OVER_PI = (REAL)-0.45158270528945486473;
  dq[0] = dq[1] = dq[2] = 0;
  *dx = 0;
  switch (kind) {
    case PM4_T_POTENTIAL:
      *dx = 1;
      return x;
    case PM4_T_NORMAL: {
      REAL mu = q[0], s = q[1];
      REAL w = x / s - mu / s; /* tfd.Normal divides before subtracting */
      *dx = -w / s;
      dq[0] = w / s;
      dq[1] = (w * w - 1) / s;
      return (REAL)-0.5 * w * w - HALF_LOG_2PI - RLOG(s);
    }
    case PM4_T_HALFNORMAL: {
      REAL s = q[0];
      REAL w = x / s;
      *dx = -w / s;
      dq[0] = (w * w - 1) / s;
      if (x < 0) return -(REAL)INFINITY;
      return (REAL)-0.5 * w * w - HALF_LOG_HALF_PI - RLOG(s);
    }
    case PM4_T_HALFCAUCHY: {
      REAL s = q[0];
      REAL w = x / s;
      REAL den = 1 + w * w;
      *dx = -2 * w / (s * den);
      dq[0] = (2 * w * w / den - 1) / s;
      if (x < 0) return -(REAL)INFINITY;
      return LOG_2_OVER_PI - RLOG(s) - RLOG1P(w * w);
    }
    case PM4_T_BERNOULLI: {
      REAL p = q[0];
      /* xlogy(x, p) + xlog1py(1 - x, -p) */
      REAL a = (x == 0) ? 0 : x * RLOG(p);
      REAL b = (1 - x == 0) ? 0 : (1 - x) * RLOG1P(-p);
      dq[0] = ((x == 0) ? 0 : x / p) - ((1 - x == 0) ? 0 : (1 - x) / (1 - p));
      *dx = 0; /* a free Bernoulli value is a frozen coordinate of the gradient-based blocks (compound step) */
      return a + b;
    }
    case PM4_T_POISSON: {
      REAL lam = q[0];
      REAL a = (x == 0) ? 0 : x * RLOG(lam);
      dq[0] = ((x == 0) ? 0 : x / lam) - 1;
      *dx = 0; /* free Poisson values are rejected by the host (discrete + gradient sampler) */
      return a - lam - RLGAMMA(x + 1);
    }
  }
  return (REAL)NAN;
}

static inline REAL NAME(digamma)(REAL x) {
  REAL r = 0;
  while (x < 6) { r -= 1 / x; x += 1; }
  REAL f = 1 / (x * x);
  return r + RLOG(x) - (REAL)0.5 / x -
         f * ((REAL)(1.0 / 12) - f * ((REAL)(1.0 / 120) - f * ((REAL)(1.0 / 252) - f * ((REAL)(1.0 / 240)))));
}

static inline REAL NAME(sigmoid)(REAL v) { return 1 / (1 + REXP(-v)); }
static inline REAL NAME(softplus)(REAL v) { return v > 0 ? v + RLOG1P(REXP(-v)) : RLOG1P(REXP(v)); }

/* ----------------------------------------------------------------------------------------
 * Tape VM: one term, all n elements.  Follows SamplingState.collect_log_prob
 * (/root/reference/pymc4/flow/executor.py:171-187): sum over distributions of
 * reduce_sum(dist.log_prob(value)) + sum over potentials of reduce_sum(value).
 * Adjoint sweep = reverse-mode autodiff of the same tape (what tf.GradientTape does for
 * tfp.math.value_and_gradient in the leapfrog integrator).
 * -------------------------------------------------------------------------------------- */
#define PM4O_MAXREG 256

static void NAME(tape_forward)(const pm4_ir_t* ir, int ob, int oe, int i, const REAL* x,
                               const REAL* z, REAL* v) {
  for (int k = ob; k < oe; ++k) {
    const pm4_op_t* o = &ir->ops[k];
    REAL a = o->a >= 0 ? v[o->a] : 0, b = o->b >= 0 ? v[o->b] : 0, r = 0;
    switch (o->opcode) {
      case PM4_OP_LD_CONST: r = (REAL)o->imm; break;
      case PM4_OP_LD_DATA: r = (REAL)ir->dataf[o->blob + (o->mode == PM4_MODE_ELEM ? i : 0)]; break;
      case PM4_OP_LD_X:
      case PM4_OP_LD_Z: {
        int64_t pos = o->base;
        if (o->mode == PM4_MODE_ELEM) pos += i;
        else if (o->mode == PM4_MODE_GATHER) pos += ir->datai[o->blob + i];
        r = (o->opcode == PM4_OP_LD_X) ? x[pos] : z[pos];
      } break;
      case PM4_OP_ADD: r = a + b; break;
      case PM4_OP_SUB: r = a - b; break;
      case PM4_OP_MUL: r = a * b; break;
      case PM4_OP_DIV: r = a / b; break;
      case PM4_OP_POW: r = RPOW(a, b); break;
      case PM4_OP_NEG: r = -a; break;
      case PM4_OP_EXP: r = REXP(a); break;
      case PM4_OP_LOG: r = RLOG(a); break;
      case PM4_OP_LOG1P: r = RLOG1P(a); break;
      case PM4_OP_SQRT: r = RSQRT(a); break;
      case PM4_OP_SQUARE: r = a * a; break;
      case PM4_OP_SIGMOID: r = NAME(sigmoid)(a); break;
      case PM4_OP_SOFTPLUS: r = NAME(softplus)(a); break;
      case PM4_OP_TANH: r = RTANH(a); break;
      case PM4_OP_ABS: r = RABS(a); break;
      case PM4_OP_RECIP: r = 1 / a; break;
      case PM4_OP_LGAMMA: r = RLGAMMA(a); break;
    }
    v[o->dst] = r;
  }
}

static void NAME(tape_backward)(const pm4_ir_t* ir, int ob, int oe, int i, const REAL* v, REAL* adj,
                                REAL* gx, REAL* gz) {
  for (int k = oe - 1; k >= ob; --k) {
    const pm4_op_t* o = &ir->ops[k];
    REAL g = adj[o->dst];
    if (g == 0) continue;
    REAL a = o->a >= 0 ? v[o->a] : 0, b = o->b >= 0 ? v[o->b] : 0, r = v[o->dst];
    switch (o->opcode) {
      case PM4_OP_LD_X:
      case PM4_OP_LD_Z: {
        int64_t pos = o->base;
        if (o->mode == PM4_MODE_ELEM) pos += i;
        else if (o->mode == PM4_MODE_GATHER) pos += ir->datai[o->blob + i];
        if (o->opcode == PM4_OP_LD_X) gx[pos] += g; else gz[pos] += g;
      } break;
      case PM4_OP_ADD: adj[o->a] += g; adj[o->b] += g; break;
      case PM4_OP_SUB: adj[o->a] += g; adj[o->b] -= g; break;
      case PM4_OP_MUL: adj[o->a] += g * b; adj[o->b] += g * a; break;
      case PM4_OP_DIV: adj[o->a] += g / b; adj[o->b] -= g * r / b; break;
      case PM4_OP_POW: adj[o->a] += g * b * RPOW(a, b - 1); if (a > 0) adj[o->b] += g * r * RLOG(a); break;
      case PM4_OP_NEG: adj[o->a] -= g; break;
      case PM4_OP_EXP: adj[o->a] += g * r; break;
      case PM4_OP_LOG: adj[o->a] += g / a; break;
      case PM4_OP_LOG1P: adj[o->a] += g / (1 + a); break;
      case PM4_OP_SQRT: adj[o->a] += g / (2 * r); break;
      case PM4_OP_SQUARE: adj[o->a] += 2 * g * a; break;
      case PM4_OP_SIGMOID: adj[o->a] += g * r * (1 - r); break;
      case PM4_OP_SOFTPLUS: adj[o->a] += g * NAME(sigmoid)(a); break;
      case PM4_OP_TANH: adj[o->a] += g * (1 - r * r); break;
      case PM4_OP_ABS: adj[o->a] += (a > 0) ? g : ((a < 0) ? -g : 0); break;
      case PM4_OP_RECIP: adj[o->a] -= g * r * r; break;
      case PM4_OP_LGAMMA: adj[o->a] += g * NAME(digamma)(a); break;
      default: break;
    }
  }
}

/* workspace for one chain */
typedef struct NAME(ws) {
  int D;
  REAL *x, *gx, *gz;
} NAME(ws_t);

/* x = shift + scale * f(theta): inverse of the unconstraining transform
 * (/root/reference/pymc4/flow/transformed_executor.py:73-75; transforms.py:116-181) */
static void NAME(constrain)(const pm4_ir_t* ir, const REAL* theta, REAL* x) {
  for (int r = 0; r < ir->n_rv; ++r) {
    const pm4_rv_t* rv = &ir->rvs[r];
    for (int j = rv->offset; j < rv->offset + rv->size; ++j) {
      if (rv->transform == PM4_TR_EXP) x[j] = (REAL)rv->shift + (REAL)rv->scale * REXP(theta[j]);
      else if (rv->transform == PM4_TR_SIGMOID) x[j] = (REAL)rv->shift + (REAL)rv->scale * NAME(sigmoid)(theta[j]);
      else x[j] = theta[j];
    }
  }
}

/* element i of a dense Bernoulli-logit term: eta = X[i, :] . beta, p = sigmoid(eta), returns log_prob(y_i) */
static REAL NAME(glm_elem)(const pm4_ir_t* ir, const pm4_term_t* tm, const REAL* x, int i, REAL* eta_out, REAL* p_out) {
  REAL eta = 0;
  for (int q = 0; q < tm->glm_p; ++q) eta += (REAL)ir->dataf[tm->glm_x + (int64_t)i * tm->glm_p + q] * x[tm->glm_base + q];
  const REAL p = NAME(sigmoid)(eta);
  const REAL yv = (REAL)ir->dataf[tm->glm_y + i];
  REAL dx, dq[3];
  const REAL q1[3] = {p, 0, 0};
  *eta_out = eta;
  *p_out = p;
  return NAME(lpdf)(PM4_T_BERNOULLI, yv, q1, &dx, dq);
}

/* Marginal GP likelihood (PM4_T_MVN_EXPQUAD): y ~ MvNormal(0, K), K_ij = a^2 exp(-|X_i - X_j|^2 / (2 l^2)) +
 * (noise + jitter) [i == j]  (/root/reference/pymc4/gp/cov.py:399-475 -> tfp ExponentiatedQuadratic.matrix;
 * distributions/multivariate.py:159-199 -> tfd.MultivariateNormalFullCovariance.log_prob =
 * -1/2 y^T K^-1 y - sum log diag chol(K) - n/2 log 2 pi).  The gradient with respect to the three scalars is the
 * reverse-mode result of that expression, 1/2 tr((alpha alpha^T - K^-1) dK/dp) with alpha = K^-1 y.  The term is
 * factorised in float64 whatever REAL is (as on the device); gx may be NULL.  Returns NAN when K is not
 * positive definite. */
static double NAME(gp_term)(const pm4_ir_t* ir, const pm4_term_t* tm, const REAL* x, REAL* gx) {
  const int n = tm->glm_base, d = tm->glm_p;
  const double *X = ir->dataf + tm->glm_x, *y = ir->dataf + tm->glm_y, *par = ir->dataf + tm->gp_par;
  const int el = (int)par[0], ea = (int)par[1], en = (int)par[2], sq = par[6] != 0.0;
  const double l = el >= 0 ? (double)x[el] : par[3], a = ea >= 0 ? (double)x[ea] : par[4];
  const double v = en >= 0 ? (double)x[en] : par[5];
  const double dg = (sq ? v * v : v) + par[7];
  double* K = (double*)malloc(sizeof(double) * (size_t)n * n);
  double* Z = gx ? (double*)calloc((size_t)n * n, sizeof(double)) : NULL;
  double* w = (double*)malloc(sizeof(double) * (size_t)n * 2);
  double* alpha = w + n;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double d2 = 0;
      for (int f = 0; f < d; ++f) { const double df = X[(size_t)i * d + f] - X[(size_t)j * d + f]; d2 += df * df; }
      K[(size_t)i * n + j] = a * a * exp(-d2 / (2.0 * l * l)) + (i == j ? dg : 0.0);
    }
  double lp = NAN, logdet = 0;
  int ok = 1;
  for (int j = 0; j < n && ok; ++j) { /* Cholesky, lower, in place */
    double s = K[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) s -= K[(size_t)j * n + k] * K[(size_t)j * n + k];
    if (!(s > 0)) { ok = 0; break; }
    const double ljj = sqrt(s);
    K[(size_t)j * n + j] = ljj;
    logdet += log(ljj);
    for (int i = j + 1; i < n; ++i) {
      double t = K[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) t -= K[(size_t)i * n + k] * K[(size_t)j * n + k];
      K[(size_t)i * n + j] = t / ljj;
    }
  }
  if (ok) {
    double quad = 0;
    for (int i = 0; i < n; ++i) { /* w = L^-1 y */
      double t = y[i];
      for (int k = 0; k < i; ++k) t -= K[(size_t)i * n + k] * w[k];
      w[i] = t / K[(size_t)i * n + i];
      quad += w[i] * w[i];
    }
    lp = -0.5 * quad - logdet - 0.5 * n * 1.8378770664093454835606594728112;
    if (gx) {
      /* Z = L^-1 row by row (contiguous inner loops): z_i = (e_i - sum_{k < i} L[i, k] z_k) / L[i, i] */
      for (int i = 0; i < n; ++i) {
        double* zi = Z + (size_t)i * n;
        zi[i] = 1.0;
        for (int k = 0; k < i; ++k) {
          const double lik = K[(size_t)i * n + k];
          const double* zk = Z + (size_t)k * n;
          for (int j = 0; j <= k; ++j) zi[j] -= lik * zk[j];
        }
        const double r = 1.0 / K[(size_t)i * n + i];
        for (int j = 0; j <= i; ++j) zi[j] *= r;
      }
      for (int i = 0; i < n; ++i) alpha[i] = 0; /* alpha = L^-T w */
      for (int k = 0; k < n; ++k)
        for (int i = 0; i <= k; ++i) alpha[i] += Z[(size_t)k * n + i] * w[k];
      /* K^-1 = Z^T Z accumulated into the (now free) lower triangle of K, one rank-1 update per row of Z */
      for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) K[(size_t)i * n + j] = 0;
      for (int k = 0; k < n; ++k) {
        const double* zk = Z + (size_t)k * n;
        for (int i = 0; i <= k; ++i) {
          const double zki = zk[i];
          double* gi = K + (size_t)i * n;
          for (int j = 0; j <= i; ++j) gi[j] += zki * zk[j];
        }
      }
      double s1 = 0, s2 = 0, s3 = 0;
      for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
          const double W = alpha[i] * alpha[j] - K[(size_t)i * n + j];
          double d2 = 0;
          for (int f = 0; f < d; ++f) { const double df = X[(size_t)i * d + f] - X[(size_t)j * d + f]; d2 += df * df; }
          const double kse = a * a * exp(-d2 / (2.0 * l * l)), wt = i == j ? 1.0 : 2.0;
          s1 += wt * W * kse;
          s2 += wt * W * kse * d2;
          if (i == j) s3 += W;
        }
      if (el >= 0) gx[el] += (REAL)(0.5 * s2 / (l * l * l));
      if (ea >= 0) gx[ea] += (REAL)(s1 / a);
      if (en >= 0) gx[en] += (REAL)(0.5 * s3 * (sq ? 2.0 * v : 1.0));
    }
  }
  free(K); free(Z); free(w);
  return lp;
}

/* joint log-density and gradient in unconstrained space for ONE chain */
static REAL NAME(logp_grad1)(const pm4_ir_t* ir, NAME(ws_t)* ws, const REAL* theta, REAL* grad) {
  const int D = ir->D;
  REAL *x = ws->x, *gx = ws->gx, *gz = ws->gz;
  NAME(constrain)(ir, theta, x);
  for (int j = 0; j < D; ++j) gx[j] = gz[j] = 0;
  REAL v[PM4O_MAXREG], adj[PM4O_MAXREG];
  REAL lp = 0;
  for (int t = 0; t < ir->n_terms; ++t) {
    const pm4_term_t* tm = &ir->terms[t];
    REAL term_sum = 0;
    if (tm->kind == PM4_T_MVN_EXPQUAD) {
      lp += (REAL)NAME(gp_term)(ir, tm, x, grad ? gx : NULL);
      continue;
    }
    if (tm->kind == PM4_T_BERNOULLI_LOGIT_GLM) {
      /* tfd.Bernoulli(probs=sigmoid(X beta)).log_prob(y), summed: xlogy(y, p) + xlog1py(1 - y, -p) */
      for (int i = 0; i < tm->n; ++i) {
        REAL eta, p;
        term_sum += NAME(glm_elem)(ir, tm, x, i, &eta, &p);
        if (grad) {
          const REAL yv = (REAL)ir->dataf[tm->glm_y + i];
          const REAL r = yv - p; /* d/d eta of the Bernoulli-logit log-likelihood */
          for (int q = 0; q < tm->glm_p; ++q) gx[tm->glm_base + q] += r * (REAL)ir->dataf[tm->glm_x + (int64_t)i * tm->glm_p + q];
        }
      }
      lp += term_sum;
      continue;
    }
    for (int i = 0; i < tm->n; ++i) {
      NAME(tape_forward)(ir, tm->op_begin, tm->op_end, i, x, theta, v);
      REAL q[3] = {0, 0, 0}, dq[3], dx;
      for (int k = 0; k < 3; ++k) if (tm->param_reg[k] >= 0) q[k] = v[tm->param_reg[k]];
      term_sum += NAME(lpdf)(tm->kind, v[tm->value_reg], q, &dx, dq);
      if (grad) {
        for (int k = 0; k < tm->n_regs; ++k) adj[k] = 0;
        adj[tm->value_reg] += dx;
        for (int k = 0; k < 3; ++k) if (tm->param_reg[k] >= 0) adj[tm->param_reg[k]] += dq[k];
        NAME(tape_backward)(ir, tm->op_begin, tm->op_end, i, v, adj, gx, gz);
      }
    }
    lp += term_sum; /* reduce_sum per distribution, then sum over distributions */
  }
  if (grad) {
    for (int r = 0; r < ir->n_rv; ++r) {
      const pm4_rv_t* rv = &ir->rvs[r];
      for (int j = rv->offset; j < rv->offset + rv->size; ++j) {
        REAL dxdz = 1;
        if (rv->transform == PM4_TR_EXP) dxdz = x[j] - (REAL)rv->shift;
        else if (rv->transform == PM4_TR_SIGMOID) {
          REAL s = (x[j] - (REAL)rv->shift) / (REAL)rv->scale;
          dxdz = (REAL)rv->scale * s * (1 - s);
        }
        grad[j] = gx[j] * dxdz + gz[j];
      }
    }
  }
  return lp;
}

/* pointwise log-likelihood of ONE state: dist.log_prob(observed) of every observed variable, no
 * reduction (/root/reference/pymc4/mcmc/samplers.py:832-854) */
static void NAME(loglik1)(const pm4_ir_t* ir, NAME(ws_t)* ws, const REAL* theta, REAL* out) {
  REAL* x = ws->x;
  NAME(constrain)(ir, theta, x);
  REAL v[PM4O_MAXREG];
  for (int t = 0; t < ir->n_terms; ++t) {
    const pm4_term_t* tm = &ir->terms[t];
    if (!tm->observed) continue;
    if (tm->kind == PM4_T_MVN_EXPQUAD) { /* one density value per state */
      out[tm->ll_offset] = (REAL)NAME(gp_term)(ir, tm, x, NULL);
      continue;
    }
    for (int i = 0; i < tm->n; ++i) {
      if (tm->kind == PM4_T_BERNOULLI_LOGIT_GLM) {
        REAL eta, p;
        out[tm->ll_offset + i] = NAME(glm_elem)(ir, tm, x, i, &eta, &p);
        continue;
      }
      NAME(tape_forward)(ir, tm->op_begin, tm->op_end, i, x, theta, v);
      REAL q[3] = {0, 0, 0}, dq[3], dx;
      for (int k = 0; k < 3; ++k) if (tm->param_reg[k] >= 0) q[k] = v[tm->param_reg[k]];
      out[tm->ll_offset + i] = NAME(lpdf)(tm->kind, v[tm->value_reg], q, &dx, dq);
    }
  }
}

static NAME(ws_t) NAME(ws_new)(int D) {
  NAME(ws_t) w;
  w.D = D;
  w.x = (REAL*)calloc((size_t)(3 * D + 3), sizeof(REAL));
  w.gx = w.x + D;
  w.gz = w.gx + D;
  return w;
}
static void NAME(ws_free)(NAME(ws_t)* w) { free(w->x); }

int NAME(pm4o_log_likelihood)(const pm4_ir_t* ir, int64_t M, int ll_row, const REAL* theta, REAL* out) {
  if (!ir || !theta || !out) return PM4_EINVAL;
  for (int t = 0; t < ir->n_terms; ++t) if (ir->terms[t].n_regs > PM4O_MAXREG) return PM4_EINVAL;
  NAME(ws_t) ws = NAME(ws_new)(ir->D);
  for (int64_t m = 0; m < M; ++m) NAME(loglik1)(ir, &ws, theta + m * ir->D, out + m * ll_row);
  NAME(ws_free)(&ws);
  return PM4_OK;
}

int NAME(pm4o_logp_grad)(const pm4_ir_t* ir, int C, const REAL* theta, REAL* lp, REAL* grad) {
  if (!ir || !theta || !lp) return PM4_EINVAL;
  for (int t = 0; t < ir->n_terms; ++t) if (ir->terms[t].n_regs > PM4O_MAXREG) return PM4_EINVAL;
  const int D = ir->D;
#pragma omp parallel
  {
    NAME(ws_t) ws = NAME(ws_new)(D);
#pragma omp for schedule(static)
    for (int c = 0; c < C; ++c)
      lp[c] = NAME(logp_grad1)(ir, &ws, theta + (size_t)c * D, grad ? grad + (size_t)c * D : NULL);
    NAME(ws_free)(&ws);
  }
  return PM4_OK;
}

int NAME(pm4o_deterministics)(const pm4_ir_t* ir, int64_t M, const REAL* theta, REAL* out) {
  const int D = ir->D;
#pragma omp parallel
  {
    REAL* x = (REAL*)malloc(sizeof(REAL) * (size_t)(D + 1));
    REAL v[PM4O_MAXREG];
#pragma omp for schedule(static)
    for (int64_t m = 0; m < M; ++m) {
      const REAL* th = theta + (size_t)m * D;
      NAME(constrain)(ir, th, x);
      for (int d = 0; d < ir->n_dets; ++d) {
        const pm4_det_t* dt = &ir->dets[d];
        for (int i = 0; i < dt->n; ++i) {
          NAME(tape_forward)(ir, dt->op_begin, dt->op_end, i, x, th, v);
          out[(size_t)m * ir->det_row + dt->out_offset + i] = v[dt->value_reg];
        }
      }
    }
    free(x);
  }
  return PM4_OK;
}

/* ----------------------------------------------------------------------------------------
 * Random numbers: Philox4x32-10 (Salmon et al., SC'11), counter = (chain, transition, slot,
 * purpose), key = seed.  The CUDA kernels use the identical stream (csrc/pm4_philox.cuh), which
 * is what makes the sampler host-replayable.  TFP uses TF stateless RNG with split seeds
 * (SURVEY App. A.1 "RNG consumption is positional"); the counter layout below reproduces that
 * positional structure: one momentum draw per transition, one direction bit + one merge uniform
 * per doubling, one uniform per leaf.
 * -------------------------------------------------------------------------------------- */
static inline REAL NAME(u01)(const uint32_t* r) {
#if REAL_IS_DOUBLE
  return (REAL)((((uint64_t)(r[0] >> 5)) * 67108864.0 + (double)(r[1] >> 6)) * (1.0 / 9007199254740992.0));
#else
  return (REAL)(r[0] >> 8) * (REAL)(1.0 / 16777216.0);
#endif
}

static inline void NAME(box_muller)(uint32_t xa, uint32_t xb, REAL* n0, REAL* n1) {
#if REAL_IS_DOUBLE
  REAL u1 = ((REAL)xa + 1.0) * (1.0 / 4294967296.0);
  REAL u2 = (REAL)xb * (1.0 / 4294967296.0);
#else
  REAL u1 = (REAL)((xa >> 8) + 1u) * (REAL)(1.0 / 16777216.0);
  REAL u2 = (REAL)(xb >> 8) * (REAL)(1.0 / 16777216.0);
#endif
  REAL rad = RSQRT((REAL)-2 * RLOG(u1));
  REAL ang = (REAL)6.283185307179586476925 * u2;
  *n0 = rad * RCOS(ang);
  *n1 = rad * RSIN(ang);
}

/* one draw of a variable given its parameters: posterior predictive
 * (/root/reference/pymc4/forward_sampling.py:230-427).  Same algorithm and Philox counters as
 * csrc/pm4_philox.cuh: draw_value. */
static REAL NAME(draw_value)(int kind, REAL p0, REAL p1, uint32_t c0, uint32_t c1, uint32_t c2, uint64_t seed) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t ctr[4] = {c0, c1, c2, PM4_RNG_PREDICTIVE}, u[4];
  pm4o_philox4x32_10(ctr, k0, k1, u);
  if (kind == PM4_T_NORMAL || kind == PM4_T_HALFNORMAL) {
    REAL n0, n1;
    NAME(box_muller)(u[0], u[1], &n0, &n1);
    return kind == PM4_T_NORMAL ? p0 + p1 * n0 : RABS(p0 * n0);
  }
  if (kind == PM4_T_HALFCAUCHY) return p0 * RTAN((REAL)1.57079632679489661923 * NAME(u01)(u));
  if (kind == PM4_T_BERNOULLI) return NAME(u01)(u) < p0 ? (REAL)1 : (REAL)0;
  if (kind == PM4_T_POISSON) {
    const REAL lam = p0;
    if (lam < (REAL)10) {
      const REAL uu = NAME(u01)(u);
      REAL p = REXP(-lam), F = p;
      int k = 0;
      while (uu > F && k < 1000) { k += 1; p *= lam / (REAL)k; F += p; }
      return (REAL)k;
    }
    const REAL slam = RSQRT(lam), loglam = RLOG(lam), b = (REAL)0.931 + (REAL)2.53 * slam;
    const REAL a = (REAL)-0.059 + (REAL)0.02483 * b, inv_alpha = (REAL)1.1239 + (REAL)1.1328 / (b - (REAL)3.4);
    const REAL vr = (REAL)0.9277 - (REAL)3.6224 / (b - (REAL)2);
    for (uint32_t it = 0; it < 64u; ++it) {
      uint32_t w[4];
      if (it == 0) memcpy(w, u, sizeof w);
      else { ctr[3] = PM4_RNG_PREDICTIVE | (it << 8); pm4o_philox4x32_10(ctr, k0, k1, w); }
      const REAL U = NAME(u01)(w) - (REAL)0.5, V = NAME(u01)(w + 2);
      const REAL us = (REAL)0.5 - RABS(U);
      const REAL k = RFLOOR(((REAL)2 * a / us + b) * U + lam + (REAL)0.43);
      if (us >= (REAL)0.07 && V <= vr) return k;
      if (k < 0 || (us < (REAL)0.013 && V > us)) continue;
      if (RLOG(V) + RLOG(inv_alpha) - RLOG(a / (us * us) + b) <= -lam + k * loglam - RLGAMMA(k + 1)) return k;
    }
    return RFLOOR(lam + (REAL)0.5);
  }
  return (REAL)NAN;
}

int NAME(pm4o_posterior_predictive)(const pm4_ir_t* ir, int64_t M, int ll_row, uint64_t seed, const REAL* theta,
                                    REAL* out) {
  if (!ir || !theta || !out) return PM4_EINVAL;
  for (int t = 0; t < ir->n_terms; ++t) if (ir->terms[t].n_regs > PM4O_MAXREG) return PM4_EINVAL;
  const int D = ir->D;
  REAL* x = (REAL*)malloc(sizeof(REAL) * (size_t)(D + 1));
  REAL v[PM4O_MAXREG];
  for (int64_t m = 0; m < M; ++m) {
    const REAL* th = theta + (size_t)m * D;
    NAME(constrain)(ir, th, x);
    for (int t = 0; t < ir->n_terms; ++t) {
      const pm4_term_t* tm = &ir->terms[t];
      if (!tm->observed) continue;
      if (tm->kind == PM4_T_MVN_EXPQUAD) continue; /* a vector per state: pm4o_gp_predictive */
      for (int i = 0; i < tm->n; ++i) {
        REAL q[3] = {0, 0, 0};
        int kind = tm->kind;
        if (kind == PM4_T_BERNOULLI_LOGIT_GLM) {
          REAL eta;
          NAME(glm_elem)(ir, tm, x, i, &eta, &q[0]);
          kind = PM4_T_BERNOULLI;
        } else {
          NAME(tape_forward)(ir, tm->op_begin, tm->op_end, i, x, th, v);
          for (int k = 0; k < 3; ++k) if (tm->param_reg[k] >= 0) q[k] = v[tm->param_reg[k]];
        }
        out[(size_t)m * ll_row + tm->ll_offset + i] =
            NAME(draw_value)(kind, q[0], q[1], (uint32_t)m, (uint32_t)((uint64_t)m >> 32),
                             (uint32_t)(tm->ll_offset + i), seed);
      }
    }
  }
  free(x);
  return PM4_OK;
}

/* Prior predictive: ancestral sampling of the whole model
 * (/root/reference/pymc4/forward_sampling.py:26-227: evaluate_model with no values given, once per
 * sample).  Free variables are drawn in model order from their distributions given the already
 * drawn ones (x holds CONSTRAINED values), then the observed variables; Philox element ids:
 * ll_row + position for free variables, row position for observed ones. */
static int NAME(term_defines)(const pm4_ir_t* ir, const pm4_term_t* tm) {
  if (tm->observed || tm->kind == PM4_T_POTENTIAL || tm->kind == PM4_T_BERNOULLI_LOGIT_GLM) return 0;
  const pm4_op_t* o = &ir->ops[tm->op_begin + tm->value_reg];
  return o->opcode == PM4_OP_LD_X && (o->mode == PM4_MODE_ELEM || o->mode == PM4_MODE_SCALAR);
}

int NAME(pm4o_prior_predictive)(const pm4_ir_t* ir, int64_t M, int ll_row, uint64_t seed, REAL* x_out, REAL* obs_out) {
  if (!ir || !x_out) return PM4_EINVAL;
  for (int t = 0; t < ir->n_terms; ++t) if (ir->terms[t].n_regs > PM4O_MAXREG) return PM4_EINVAL;
  const int D = ir->D;
  REAL v[PM4O_MAXREG];
  for (int64_t m = 0; m < M; ++m) {
    REAL* x = x_out + (size_t)m * D;
    for (int j = 0; j < D; ++j) x[j] = 0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int t = 0; t < ir->n_terms; ++t) {
        const pm4_term_t* tm = &ir->terms[t];
        if (tm->kind == PM4_T_MVN_EXPQUAD) continue; /* a vector per sample: pm4o_gp_predictive */
        const int defines = NAME(term_defines)(ir, tm);
        if (pass == 0 ? !defines : !(tm->observed && obs_out)) continue;
        const pm4_op_t* vo = &ir->ops[tm->op_begin + (tm->value_reg >= 0 ? tm->value_reg : 0)];
        for (int i = 0; i < tm->n; ++i) {
          REAL q[3] = {0, 0, 0};
          if (tm->kind == PM4_T_BERNOULLI_LOGIT_GLM) {
            REAL eta;
            NAME(glm_elem)(ir, tm, x, i, &eta, &q[0]);
            obs_out[(size_t)m * ll_row + tm->ll_offset + i] =
                NAME(draw_value)(PM4_T_BERNOULLI, q[0], q[1], (uint32_t)m, (uint32_t)((uint64_t)m >> 32),
                                 (uint32_t)(tm->ll_offset + i), seed);
            continue;
          }
          NAME(tape_forward)(ir, tm->op_begin, tm->op_end, i, x, x, v);
          for (int k = 0; k < 3; ++k) if (tm->param_reg[k] >= 0) q[k] = v[tm->param_reg[k]];
          if (pass == 0) {
            const int pos = vo->base + (vo->mode == PM4_MODE_ELEM ? i : 0);
            x[pos] = NAME(draw_value)(tm->kind, q[0], q[1], (uint32_t)m, (uint32_t)((uint64_t)m >> 32),
                                      (uint32_t)(ll_row + pos), seed);
          } else {
            obs_out[(size_t)m * ll_row + tm->ll_offset + i] =
                NAME(draw_value)(tm->kind, q[0], q[1], (uint32_t)m, (uint32_t)((uint64_t)m >> 32),
                                 (uint32_t)(tm->ll_offset + i), seed);
          }
        }
      }
    }
  }
  return PM4_OK;
}

/* momentum for every dimension of one chain at one transition; dimension j is owned by
 * lane j%32, register j/32 of the chain's warp and four registers share one Philox call */
static void NAME(draw_momentum)(uint64_t seed, uint32_t chain, uint32_t t, int D, REAL* p) {
  for (int j = 0; j < D; ++j) {
    int lane = j & 31, r = j >> 5, q = r >> 2, e = r & 3;
    uint32_t ctr[4] = {chain, t, (uint32_t)(q * 32 + lane), PM4_RNG_MOMENTUM}, out[4];
    pm4o_philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32), out);
    REAL n[4];
    NAME(box_muller)(out[0], out[1], &n[0], &n[1]);
    NAME(box_muller)(out[2], out[3], &n[2], &n[3]);
    p[j] = n[e];
  }
}

/* a uniform in [0,1): words (2*pair, 2*pair+1) of the Philox block at `slot` */
static inline REAL NAME(draw_u)(uint64_t seed, uint32_t chain, uint32_t t, uint32_t slot, uint32_t purpose, int pair) {
  uint32_t ctr[4] = {chain, t, slot, purpose}, out[4];
  pm4o_philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32), out);
  return NAME(u01)(out + 2 * pair);
}

/* the uniform of leaf i of the subtree at `depth`: two leaves share one Philox block */
static inline REAL NAME(draw_leaf_u)(uint64_t seed, uint32_t chain, uint32_t t, int depth, int i) {
  return NAME(draw_u)(seed, chain, t, ((uint32_t)depth << 20) | ((uint32_t)i >> 1), PM4_RNG_LEAF, i & 1);
}

/* direction bits of all doublings of one transition come from one Philox word */
static inline int NAME(draw_direction)(uint64_t seed, uint32_t chain, uint32_t t, int depth) {
  uint32_t ctr[4] = {chain, t, 0u, PM4_RNG_DIRECTION}, out[4];
  pm4o_philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32), out);
  return (int)((out[0] >> depth) & 1u);
}

/* ----------------------------------------------------------------------------------------
 * Leapfrog: tfp SimpleLeapfrogIntegrator: half kick, L x (drift, full kick),
 * half un-kick.  eps[j] = signed step * per-dimension scale.
 * -------------------------------------------------------------------------------------- */
static void NAME(leapfrog)(const pm4_ir_t* ir, NAME(ws_t)* ws, int L, REAL eps, const REAL* scale,
                           REAL* theta, REAL* p, REAL* grad, REAL* lp) {
  const int D = ir->D;
  for (int j = 0; j < D; ++j) {
    REAL e = scale ? eps * scale[j] : eps;
    p[j] = p[j] + ((REAL)0.5 * e) * grad[j];
  }
  for (int l = 0; l < L; ++l) {
    for (int j = 0; j < D; ++j) {
      REAL e = scale ? eps * scale[j] : eps;
      theta[j] = theta[j] + e * p[j];
    }
    *lp = NAME(logp_grad1)(ir, ws, theta, grad);
    for (int j = 0; j < D; ++j) {
      REAL e = scale ? eps * scale[j] : eps;
      p[j] = p[j] + e * grad[j];
    }
  }
  for (int j = 0; j < D; ++j) {
    REAL e = scale ? eps * scale[j] : eps;
    p[j] = p[j] - ((REAL)0.5 * e) * grad[j];
  }
}

static inline REAL NAME(dot)(const REAL* a, const REAL* b, int D) {
  REAL s = 0;
  for (int j = 0; j < D; ++j) s += a[j] * b[j];
  return s;
}

static inline REAL NAME(logaddexp)(REAL a, REAL b) {
  if (a == -(REAL)INFINITY) return b;
  if (b == -(REAL)INFINITY) return a;
  REAL m = a > b ? a : b;
  return m + RLOG1P(REXP(-RABS(a - b)));
}

/* ----------------------------------------------------------------------------------------
 * Dual averaging: tfp DualAveragingStepSizeAdaptation.one_step, as wired by
 * PyMC4 (/root/reference/pymc4/mcmc/samplers.py:400-405; inference/sampling.py:155-156).
 * -------------------------------------------------------------------------------------- */
typedef struct NAME(da) {
  REAL eps0, eps, error_sum, log_avg, log_shrink_target;
  int t0; /* first transition of the current adaptation window (mass-matrix windows restart the recursion) */
} NAME(da_t);

static void NAME(da_init)(NAME(da_t)* s, REAL eps0) {
  s->eps0 = eps0;
  s->eps = eps0;
  s->error_sum = 0;
  s->log_avg = 0;
  s->log_shrink_target = RLOG((REAL)10 * eps0);
  s->t0 = 0;
}

/* initial state of chain c (or of the pool): per-chain initial step sizes and tfp's `shrinkage_target` when given */
static void NAME(da_init_cfg)(NAME(da_t)* s, double step_size, const void* step_size_c, int c, const pm4_adapt_cfg_t* a) {
  NAME(da_init)(s, step_size_c ? ((const REAL*)step_size_c)[c] : (REAL)step_size);
  if (a->shrinkage_target > 0) s->log_shrink_target = (REAL)log(a->shrinkage_target);
}

/* restart of the recursion at transition t0 from the current step size (after a mass-matrix update) */
static void NAME(da_restart)(NAME(da_t)* s, int t0) {
  s->error_sum = 0;
  s->log_avg = 0;
  s->log_shrink_target = RLOG((REAL)10 * s->eps);
  s->t0 = t0;
}

/* Diagonal mass-matrix window (include/pm4b200.h: pm4_adapt_cfg_t.mass_windows): the step scale becomes the
 * cross-chain standard deviation of the unconstrained states divided by its geometric mean.  Returns 0 (and leaves
 * `scale` alone) when some dimension has no finite positive spread yet. */
static int NAME(mass_update)(const REAL* th, int C, int D, REAL* scale) {
  double* ls = (double*)malloc(sizeof(double) * (size_t)D);
  double mean_ls = 0;
  int ok = 1;
  for (int d = 0; d < D && ok; ++d) {
    double m = 0, v = 0;
    for (int c = 0; c < C; ++c) m += (double)th[(size_t)c * D + d];
    m /= C;
    for (int c = 0; c < C; ++c) { const double e = (double)th[(size_t)c * D + d] - m; v += e * e; }
    v /= C;
    if (!(v > 0) || !isfinite(v)) ok = 0;
    else { ls[d] = 0.5 * log(v); mean_ls += ls[d]; }
  }
  if (ok) {
    mean_ls /= D;
    for (int d = 0; d < D; ++d) scale[d] = (REAL)exp(ls[d] - mean_ls);
  }
  free(ls);
  return ok;
}

/* is transition t the first of a mass-matrix window?  w_k = num_adapt * k / (K + 1), k = 1 .. K */
static int NAME(mass_window_start)(const pm4_adapt_cfg_t* a, int t) {
  if (!a->enabled || a->mass_windows <= 0 || t <= 0 || t >= a->num_adaptation_steps) return 0;
  for (int k = 1; k <= a->mass_windows; ++k)
    if ((int)(((int64_t)a->num_adaptation_steps * k) / (a->mass_windows + 1)) == t) return 1;
  return 0;
}

/* called after transition t (0-based) with the (cross-chain or per-chain) mean accept prob */
static void NAME(da_update)(NAME(da_t)* s, const pm4_adapt_cfg_t* a, int t, REAL accept) {
  if (!a->enabled) return;
  if (a->kind == PM4_ADAPT_SIMPLE) { /* tfp SimpleStepSizeAdaptation.one_step: multiplicative, no averaging */
    const REAL f = 1 + (REAL)a->adaptation_rate;
    if (t < a->num_adaptation_steps) s->eps = accept > (REAL)a->target_accept_prob ? s->eps * f : s->eps / f;
    return;
  }
  REAL step = (REAL)(t - s->t0 + 1);
  s->error_sum += (REAL)a->target_accept_prob - accept;
  REAL log_eps = s->log_shrink_target -
                 s->error_sum * RSQRT(step) / (((REAL)a->step_count_smoothing + step) * (REAL)a->exploration_shrinkage);
  REAL eta = RPOW(step, -(REAL)a->decay_rate);
  s->log_avg = eta * log_eps + (1 - eta) * s->log_avg;
  if (t < a->num_adaptation_steps) s->eps = REXP(log_eps);
  else if (t == a->num_adaptation_steps) s->eps = REXP(s->log_avg);
}

/* ----------------------------------------------------------------------------------------
 * One NUTS transition for one chain: tfp NoUTurnSampler.one_step -> _loop_tree_doubling ->
 * _build_sub_tree -> _loop_build_sub_tree, in its per-chain-equivalent form
 * (a chain stops where TFP's mask for it turns false).
 * -------------------------------------------------------------------------------------- */
typedef struct NAME(nuts_buf) {
  REAL *p0, *cur_th, *cur_p, *cur_g, *oth_th, *oth_p, *oth_g, *cand_th, *cand_g, *sub_th, *sub_g,
      *rho, *rho_sub, *rho_prev, *tmp, *mem_p, *mem_rho;
} NAME(nuts_buf_t);

typedef struct NAME(nuts_out) {
  REAL lp, energy, log_accept;
  int leapfrogs, diverging, depth;
} NAME(nuts_out_t);

static void NAME(nuts_transition)(const pm4_ir_t* ir, NAME(ws_t)* ws, NAME(nuts_buf_t)* b,
                                  const pm4_nuts_cfg_t* cfg, uint32_t chain, uint32_t t, REAL eps,
                                  const REAL* scale, const REAL* injected_p, REAL* theta, REAL* grad,
                                  REAL* lp, NAME(nuts_out_t)* out) {
  const int D = ir->D;
  const size_t nb = sizeof(REAL) * (size_t)D;
  const REAL NEG_INF = -(REAL)INFINITY;
  if (injected_p) memcpy(b->p0, injected_p, nb);
  else NAME(draw_momentum)(cfg->seed, chain, t, D, b->p0);
  if (scale) /* step scale 0 = frozen dimension (a block of the compound step conditions on it): no momentum */
    for (int j = 0; j < D; ++j)
      if (scale[j] == 0) b->p0[j] = 0;
  const REAL H0 = *lp - (REAL)0.5 * NAME(dot)(b->p0, b->p0, D);

  /* both ends of the trajectory start at the current point; `cur` is the end being extended,
   * `oth` the far end; cur_is_right tells which is which */
  memcpy(b->cur_th, theta, nb); memcpy(b->cur_p, b->p0, nb); memcpy(b->cur_g, grad, nb);
  memcpy(b->oth_th, theta, nb); memcpy(b->oth_p, b->p0, nb); memcpy(b->oth_g, grad, nb);
  REAL cur_lp = *lp, oth_lp = *lp;
  int cur_is_right = 1;

  memcpy(b->cand_th, theta, nb); memcpy(b->cand_g, grad, nb);
  REAL cand_lp = *lp, cand_energy = H0, w = 0;
  memcpy(b->rho, b->p0, nb);
  REAL ediff_sum = 0;
  int nleap = 0, cont = 1, not_div = 1, depth = 0;

  while (depth < cfg->max_tree_depth && cont) {
    const int dir = NAME(draw_direction)(cfg->seed, chain, t, depth);
    if (dir != cur_is_right) { /* extend the other end: swap roles */
      REAL* s;
      s = b->cur_th; b->cur_th = b->oth_th; b->oth_th = s;
      s = b->cur_p; b->cur_p = b->oth_p; b->oth_p = s;
      s = b->cur_g; b->cur_g = b->oth_g; b->oth_g = s;
      REAL tl = cur_lp; cur_lp = oth_lp; oth_lp = tl;
      cur_is_right = dir;
    }
    const REAL eps_s = dir ? eps : -eps;
    const int nsteps = 1 << depth;
    for (int j = 0; j < D; ++j) b->rho_sub[j] = 0;
    REAL w_sub = NEG_INF, sub_lp = cur_lp, sub_energy = cur_lp, sub_ediff = 0;
    memcpy(b->sub_th, b->cur_th, nb); memcpy(b->sub_g, b->cur_g, nb);
    int cont_sub = 1, sub_leaps = 0;

    for (int i = 0; i < nsteps && cont_sub; ++i) {
      NAME(leapfrog)(ir, ws, 1, eps_s, scale, b->cur_th, b->cur_p, b->cur_g, &cur_lp);
      sub_leaps += 1;
      memcpy(b->rho_prev, b->rho_sub, nb);
      for (int j = 0; j < D; ++j) b->rho_sub[j] += b->cur_p[j];

      /* U-turn checks against the checkpoints that close a balanced sub-subtree at leaf i:
       * slots [popcount(i) - trailing_ones(i), popcount(i)) -- the closed form of tfp's
       * build_tree_uturn_instruction + generate_efficient_write_read_instruction tables */
      int no_uturn = 1;
      if (i & 1) {
        int pc = __builtin_popcount((unsigned)i), t1 = __builtin_ctz(~(unsigned)i);
        for (int s = pc - t1; s < pc && no_uturn; ++s) {
          const REAL* mp = b->mem_p + (size_t)s * D;
          const REAL* mr = b->mem_rho + (size_t)s * D;
          for (int j = 0; j < D; ++j) b->tmp[j] = b->rho_sub[j] - mr[j];
          no_uturn = (NAME(dot)(b->tmp, mp, D) > 0) && (NAME(dot)(b->tmp, b->cur_p, D) > 0);
        }
      } else { /* even leaves are checkpointed into slot popcount(i) */
        int s = __builtin_popcount((unsigned)i);
        memcpy(b->mem_p + (size_t)s * D, b->cur_p, nb);
        memcpy(b->mem_rho + (size_t)s * D, b->rho_prev, nb);
      }

      REAL energy = cur_lp - (REAL)0.5 * NAME(dot)(b->cur_p, b->cur_p, D);
      if (isnan(energy)) energy = NEG_INF;
      const REAL dH = energy - H0;
      const int not_divergent = (-dH < (REAL)cfg->max_energy_diff);
      const REAL w_new = NAME(logaddexp)(w_sub, dH);
      const REAL thresh = dH - w_new;
      const REAL u = RLOG1P(-NAME(draw_leaf_u)(cfg->seed, chain, t, depth, i));
      if (u <= thresh) {
        memcpy(b->sub_th, b->cur_th, nb); memcpy(b->sub_g, b->cur_g, nb);
        sub_lp = cur_lp; sub_energy = energy;
      }
      w_sub = w_new;
      const int cont_leaf = not_divergent; /* & continue_tree_previous, which is 1 here */
      not_div = not_div && not_divergent;
      if (cont_leaf) sub_ediff += REXP(dH < 0 ? dH : 0);
      cont_sub = no_uturn && cont_leaf;
    }

    /* merge the subtree into the trajectory (biased progressive sampling) */
    ediff_sum += sub_ediff;
    nleap += sub_leaps;
    const REAL tree_w = cont_sub ? w_sub : NEG_INF;
    const REAL thresh = tree_w - w;
    const REAL um = RLOG1P(-NAME(draw_u)(cfg->seed, chain, t, (uint32_t)depth, PM4_RNG_MERGE, 0));
    if ((um <= thresh) && cont_sub) {
      memcpy(b->cand_th, b->sub_th, nb); memcpy(b->cand_g, b->sub_g, nb);
      cand_lp = sub_lp; cand_energy = sub_energy;
    }
    w = NAME(logaddexp)(tree_w, w);
    for (int j = 0; j < D; ++j) b->rho[j] += b->rho_sub[j];
    const REAL *p_left = cur_is_right ? b->oth_p : b->cur_p, *p_right = cur_is_right ? b->cur_p : b->oth_p;
    const int no_uturn_traj = (NAME(dot)(b->rho, p_left, D) > 0) && (NAME(dot)(b->rho, p_right, D) > 0);
    cont = cont_sub && no_uturn_traj;
    depth += 1;
  }
  memcpy(theta, b->cand_th, nb); memcpy(grad, b->cand_g, nb);
  *lp = cand_lp;
  out->lp = cand_lp;
  out->energy = cand_energy;
  out->log_accept = RLOG(ediff_sum / (REAL)nleap);
  out->leapfrogs = nleap;
  out->diverging = !not_div;
  out->depth = depth;
}

static NAME(nuts_buf_t) NAME(nuts_buf_new)(int D, int max_depth) {
  NAME(nuts_buf_t) b;
  size_t nvec = 15 + 2 * (size_t)(max_depth + 1);
  REAL* base = (REAL*)calloc(nvec * (size_t)(D + 1), sizeof(REAL));
  REAL** f[] = {&b.p0, &b.cur_th, &b.cur_p, &b.cur_g, &b.oth_th, &b.oth_p, &b.oth_g, &b.cand_th,
                &b.cand_g, &b.sub_th, &b.sub_g, &b.rho, &b.rho_sub, &b.rho_prev, &b.tmp};
  for (int k = 0; k < 15; ++k) *f[k] = base + (size_t)k * D;
  b.mem_p = base + (size_t)15 * D;
  b.mem_rho = b.mem_p + (size_t)(max_depth + 1) * D;
  return b;
}
/* the base pointer may have been swapped between cur_* and oth_*: free the smallest */
static void NAME(nuts_buf_free)(NAME(nuts_buf_t)* b) { free(b->p0); }

/* sample_chain(num_results=S, num_burnin_steps=B) with NUTS under dual averaging
 * (/root/reference/pymc4/mcmc/samplers.py:172-189). */
int NAME(pm4o_nuts_run)(const pm4_ir_t* ir, const pm4_nuts_cfg_t* cfg, const REAL* theta0,
                        const REAL* step_scale, const REAL* momenta, REAL* states, REAL* lp_out,
                        int32_t* leapfrogs, uint8_t* diverging, REAL* energy, REAL* log_accept,
                        int32_t* depth_out, REAL* step_size_out, int64_t* total_leapfrogs, REAL* mass_scale_out) {
  const int D = ir->D, C = cfg->C, B = cfg->num_burnin, S = cfg->num_results, T = B + S;
  if (cfg->max_tree_depth < 1 || cfg->max_tree_depth > 20) return PM4_EINVAL;
  REAL* th = (REAL*)malloc(sizeof(REAL) * (size_t)C * D);
  REAL* gr = (REAL*)malloc(sizeof(REAL) * (size_t)C * D);
  REAL* lp = (REAL*)malloc(sizeof(REAL) * (size_t)C);
  REAL* acc = (REAL*)malloc(sizeof(REAL) * (size_t)C);
  int64_t* tot = (int64_t*)calloc((size_t)C, sizeof(int64_t));
  memcpy(th, theta0, sizeof(REAL) * (size_t)C * D);
  const int pooled = cfg->adapt.mode == PM4_EPS_POOLED;
  NAME(da_t)* da = (NAME(da_t)*)malloc(sizeof(NAME(da_t)) * (size_t)(pooled ? 1 : C));
  for (int c = 0; c < (pooled ? 1 : C); ++c) NAME(da_init_cfg)(&da[c], cfg->step_size, pooled ? NULL : cfg->step_size_dev, c, &cfg->adapt);

  /* the step scale in effect: the caller's (or ones when mass windows will write it), updated at window starts */
  REAL* scale_buf = NULL;
  if (cfg->adapt.enabled && cfg->adapt.mass_windows > 0) {
    scale_buf = (REAL*)malloc(sizeof(REAL) * (size_t)D);
    for (int d = 0; d < D; ++d) scale_buf[d] = step_scale ? step_scale[d] : (REAL)1;
    step_scale = scale_buf;
  }
  /* segments between mass-matrix window starts: inside a segment the chains only couple through a pooled step size */
  int* seg = (int*)malloc(sizeof(int) * (size_t)(cfg->adapt.mass_windows > 0 ? cfg->adapt.mass_windows + 3 : 3));
  int n_seg = 0;
  seg[n_seg++] = 0;
  for (int t = 1; t < T; ++t)
    if (NAME(mass_window_start)(&cfg->adapt, t)) seg[n_seg++] = t;
  seg[n_seg] = T;

#pragma omp parallel
  {
    NAME(ws_t) ws = NAME(ws_new)(D);
    NAME(nuts_buf_t) nb = NAME(nuts_buf_new)(D, cfg->max_tree_depth);
    REAL* base = nb.p0;
    /* bootstrap_results: log-density and gradient at the initial point */
#pragma omp for schedule(static)
    for (int c = 0; c < C; ++c) lp[c] = NAME(logp_grad1)(ir, &ws, th + (size_t)c * D, gr + (size_t)c * D);

    for (int sg = 0; sg < n_seg; ++sg) {
      const int ta = seg[sg], tb = seg[sg + 1];
      if (sg > 0) {
#pragma omp single
        {
          if (NAME(mass_update)(th, C, D, scale_buf))
            for (int c = 0; c < (pooled ? 1 : C); ++c) NAME(da_restart)(&da[c], ta);
        }
      }
      if (pooled) {
        for (int t = ta; t < tb; ++t) {
#pragma omp for schedule(dynamic, 1)
          for (int c = 0; c < C; ++c) {
            NAME(nuts_out_t) o;
            NAME(nuts_transition)(ir, &ws, &nb, cfg, (uint32_t)(cfg->chain_offset + c), (uint32_t)t, da[0].eps, step_scale,
                                  momenta ? momenta + ((size_t)t * C + c) * D : NULL, th + (size_t)c * D,
                                  gr + (size_t)c * D, &lp[c], &o);
            acc[c] = REXP(o.log_accept);
            tot[c] += o.leapfrogs;
            if (t >= B) {
              size_t k = (size_t)(t - B) * C + c;
              if (states) memcpy(states + k * D, th + (size_t)c * D, sizeof(REAL) * (size_t)D);
              if (lp_out) lp_out[k] = o.lp;
              if (leapfrogs) leapfrogs[k] = o.leapfrogs;
              if (diverging) diverging[k] = (uint8_t)o.diverging;
              if (energy) energy[k] = o.energy;
              if (log_accept) log_accept[k] = o.log_accept;
              if (depth_out) depth_out[k] = o.depth;
            }
          }
#pragma omp single
          {
            /* reduce_logmeanexp over all chains, then exp: the arithmetic mean accept prob */
            REAL s = 0;
            for (int c = 0; c < C; ++c) s += acc[c];
            NAME(da_update)(&da[0], &cfg->adapt, t, s / (REAL)C);
          }
        }
      } else {
#pragma omp for schedule(dynamic, 1)
        for (int c = 0; c < C; ++c) {
          for (int t = ta; t < tb; ++t) {
            NAME(nuts_out_t) o;
            NAME(nuts_transition)(ir, &ws, &nb, cfg, (uint32_t)(cfg->chain_offset + c), (uint32_t)t, da[c].eps, step_scale,
                                  momenta ? momenta + ((size_t)t * C + c) * D : NULL, th + (size_t)c * D,
                                  gr + (size_t)c * D, &lp[c], &o);
            NAME(da_update)(&da[c], &cfg->adapt, t, REXP(o.log_accept));
            tot[c] += o.leapfrogs;
            if (t >= B) {
              size_t k = (size_t)(t - B) * C + c;
              if (states) memcpy(states + k * D, th + (size_t)c * D, sizeof(REAL) * (size_t)D);
              if (lp_out) lp_out[k] = o.lp;
              if (leapfrogs) leapfrogs[k] = o.leapfrogs;
              if (diverging) diverging[k] = (uint8_t)o.diverging;
              if (energy) energy[k] = o.energy;
              if (log_accept) log_accept[k] = o.log_accept;
              if (depth_out) depth_out[k] = o.depth;
            }
          }
        }
      }
    }
    nb.p0 = base;
    NAME(nuts_buf_free)(&nb);
    NAME(ws_free)(&ws);
  }
  free(seg);
  if (mass_scale_out && scale_buf) memcpy(mass_scale_out, scale_buf, sizeof(REAL) * (size_t)D);
  free(scale_buf);
  for (int c = 0; c < C; ++c) {
    if (step_size_out) step_size_out[c] = da[pooled ? 0 : c].eps;
    if (total_leapfrogs) total_leapfrogs[c] = tot[c];
  }
  free(th); free(gr); free(lp); free(acc); free(tot); free(da);
  return PM4_OK;
}

/* sample_chain with tfp HamiltonianMonteCarlo = MetropolisHastings(UncalibratedHMC)
 * (/root/reference/pymc4/mcmc/samplers.py:265-315). */
int NAME(pm4o_hmc_run)(const pm4_ir_t* ir, const pm4_hmc_cfg_t* cfg, const REAL* theta0,
                       const REAL* step_scale, const REAL* momenta, const REAL* uniforms, REAL* states,
                       REAL* lp_out, REAL* log_accept, uint8_t* accepted, REAL* step_size_out, REAL* traj) {
  const int D = ir->D, C = cfg->C, B = cfg->num_burnin, S = cfg->num_results, T = B + S;
  const int L = cfg->num_leapfrog_steps;
  const int pooled = cfg->adapt.mode == PM4_EPS_POOLED;
  REAL* th = (REAL*)malloc(sizeof(REAL) * (size_t)C * D);
  REAL* gr = (REAL*)malloc(sizeof(REAL) * (size_t)C * D);
  REAL* lp = (REAL*)malloc(sizeof(REAL) * (size_t)C);
  REAL* acc = (REAL*)malloc(sizeof(REAL) * (size_t)C);
  memcpy(th, theta0, sizeof(REAL) * (size_t)C * D);
  NAME(da_t)* da = (NAME(da_t)*)malloc(sizeof(NAME(da_t)) * (size_t)(pooled ? 1 : C));
  for (int c = 0; c < (pooled ? 1 : C); ++c) NAME(da_init_cfg)(&da[c], cfg->step_size, pooled ? NULL : cfg->step_size_dev, c, &cfg->adapt);
  NAME(ws_t) ws = NAME(ws_new)(D);
  REAL* pth = (REAL*)malloc(sizeof(REAL) * (size_t)D * 3);
  REAL *pp = pth + D, *pg = pp + D;
  for (int c = 0; c < C; ++c) lp[c] = NAME(logp_grad1)(ir, &ws, th + (size_t)c * D, gr + (size_t)c * D);
  for (int t = 0; t < T; ++t) {
    for (int c = 0; c < C; ++c) {
      REAL eps = da[pooled ? 0 : c].eps;
      REAL *cth = th + (size_t)c * D, *cgr = gr + (size_t)c * D;
      if (momenta) memcpy(pp, momenta + ((size_t)t * C + c) * D, sizeof(REAL) * (size_t)D);
      else NAME(draw_momentum)(cfg->seed, (uint32_t)(cfg->chain_offset + c), (uint32_t)t, D, pp);
      const REAL k0 = (REAL)0.5 * NAME(dot)(pp, pp, D);
      memcpy(pth, cth, sizeof(REAL) * (size_t)D);
      memcpy(pg, cgr, sizeof(REAL) * (size_t)D);
      REAL plp = lp[c], lar;
      if (cfg->proposal == PM4_PROPOSAL_RANDOM_WALK) {
        /* tfp RandomWalkMetropolis with random_walk_normal_fn(scale): symmetric proposal, gradient-free
         * (/root/reference/pymc4/mcmc/samplers.py:450-478) */
        for (int j = 0; j < D; ++j) pth[j] = pth[j] + (REAL)cfg->rw_scale * pp[j];
        plp = NAME(logp_grad1)(ir, &ws, pth, pg);
        lar = plp - lp[c];
      } else {
        NAME(leapfrog)(ir, &ws, L, eps, step_scale, pth, pp, pg, &plp);
        const REAL k1 = (REAL)0.5 * NAME(dot)(pp, pp, D);
        lar = (plp - lp[c]) + (k0 - k1);
      }
      if (!isfinite(lar)) lar = -(REAL)INFINITY;
      REAL u = uniforms ? uniforms[(size_t)t * C + c]
                        : NAME(draw_u)(cfg->seed, (uint32_t)(cfg->chain_offset + c), (uint32_t)t, 0, PM4_RNG_HMC, 0);
      const int ok = RLOG(u) < lar;
      if (traj) memcpy(traj + ((size_t)t * C + c) * D, pth, sizeof(REAL) * (size_t)D);
      if (ok) {
        memcpy(cth, pth, sizeof(REAL) * (size_t)D);
        memcpy(cgr, pg, sizeof(REAL) * (size_t)D);
        lp[c] = plp;
      }
      acc[c] = REXP(lar < 0 ? lar : 0);
      if (!pooled) NAME(da_update)(&da[c], &cfg->adapt, t, acc[c]);
      if (t >= B) {
        size_t k = (size_t)(t - B) * C + c;
        if (states) memcpy(states + k * D, cth, sizeof(REAL) * (size_t)D);
        if (lp_out) lp_out[k] = lp[c];
        if (log_accept) log_accept[k] = lar;
        if (accepted) accepted[k] = (uint8_t)ok;
      }
    }
    if (pooled) {
      REAL s = 0;
      for (int c = 0; c < C; ++c) s += acc[c];
      NAME(da_update)(&da[0], &cfg->adapt, t, s / (REAL)C);
    }
  }
  for (int c = 0; c < C; ++c) if (step_size_out) step_size_out[c] = da[pooled ? 0 : c].eps;
  NAME(ws_free)(&ws);
  free(pth); free(th); free(gr); free(lp); free(acc); free(da);
  return PM4_OK;
}


/* ----------------------------------------------------------------------------------------
 * Compound step (/root/reference/pymc4/mcmc/samplers.py:490-726 CompoundStep,
 * mcmc/tf_support.py:52-163 _CompoundStepTF.one_step, distributions/state_functions.py): per transition one
 * transition of every block's kernel in order on the same chain state; a block's kernel sees the dimensions it does
 * not own (step scale 0) frozen at their current values -- _target_log_prob_fn_part_compound.  Every block has its
 * own adaptation state and its own Philox key (seed + golden ratio * block).  Mirrors pm4_compound_run
 * (csrc/pm4_core.cu); `dim_scale_dev` / `rw_kind_dev` are HOST pointers here.
 * -------------------------------------------------------------------------------------- */
static void NAME(mh_transition)(const pm4_ir_t* ir, NAME(ws_t)* ws, const pm4_compound_block_t* k, uint64_t seed,
                                uint32_t chain, uint32_t t, REAL eps, REAL* work /* 3 D */, REAL* cth, REAL* cgr, REAL* lp,
                                REAL* lar_out, int* ok_out) {
  const int D = ir->D;
  const REAL* scale = (const REAL*)k->dim_scale_dev;
  REAL *pth = work, *pp = work + D, *pg = work + 2 * D;
  NAME(draw_momentum)(seed, chain, t, D, pp);
  for (int j = 0; j < D; ++j)
    if (scale[j] == 0) pp[j] = 0;
  const REAL k0 = (REAL)0.5 * NAME(dot)(pp, pp, D);
  memcpy(pth, cth, sizeof(REAL) * (size_t)D);
  memcpy(pg, cgr, sizeof(REAL) * (size_t)D);
  REAL plp = *lp, lar;
  if (k->kind == PM4_BLOCK_RWM) {
    const int32_t* kinds = k->rw_kind_dev;
    for (int j = 0; j < D; ++j) {
      if (scale[j] == 0) continue;
      const int kd = kinds ? kinds[j] : PM4_RW_NORMAL, kind = kd & 0xff;
      const REAL z = pp[j];
      if (kind == PM4_RW_BERNOULLI) pth[j] = z > 0 ? (REAL)1 - pth[j] : pth[j];
      else if (kind == PM4_RW_CATEGORICAL) {
        const int K = kd >> 8;
        int c = (int)((REAL)K * (REAL)0.5 * (REAL)erfc((double)(-z) * 0.70710678118654752440));
        pth[j] = (REAL)(c < 0 ? 0 : (c > K - 1 ? K - 1 : c));
      } else if (kind == PM4_RW_GAUSSIAN_ROUND) pth[j] = (REAL)rint((double)(pth[j] + (REAL)k->rw_scale * z));
      else pth[j] = pth[j] + (REAL)k->rw_scale * z;
    }
    plp = NAME(logp_grad1)(ir, ws, pth, pg);
    lar = plp - *lp;
  } else {
    NAME(leapfrog)(ir, ws, k->num_leapfrog_steps, eps, scale, pth, pp, pg, &plp);
    const REAL k1 = (REAL)0.5 * NAME(dot)(pp, pp, D);
    lar = (plp - *lp) + (k0 - k1);
  }
  if (!isfinite(lar)) lar = -(REAL)INFINITY;
  const REAL u = NAME(draw_u)(seed, chain, t, 0, PM4_RNG_HMC, 0);
  const int ok = RLOG(u) < lar;
  if (ok) {
    memcpy(cth, pth, sizeof(REAL) * (size_t)D);
    memcpy(cgr, pg, sizeof(REAL) * (size_t)D);
    *lp = plp;
  }
  *lar_out = lar;
  *ok_out = ok;
}

int NAME(pm4o_compound_run)(const pm4_ir_t* ir, int C, int S, int B, uint64_t seed, int chain_offset,
                            const pm4_compound_block_t* blocks, int n_blocks, const REAL* theta0, REAL* states,
                            REAL* lp_out, REAL* step_size_out, REAL* log_accept, int32_t* leapfrogs_out) {
  const int D = ir->D, T = B + S;
  if (n_blocks < 1 || C < 1) return PM4_EINVAL;
  REAL* th = (REAL*)malloc(sizeof(REAL) * (size_t)C * D);
  REAL* gr = (REAL*)malloc(sizeof(REAL) * (size_t)C * D);
  REAL* lp = (REAL*)malloc(sizeof(REAL) * (size_t)C);
  REAL* acc = (REAL*)malloc(sizeof(REAL) * (size_t)C);
  REAL* work = (REAL*)malloc(sizeof(REAL) * (size_t)D * 3);
  memcpy(th, theta0, sizeof(REAL) * (size_t)C * D);
  NAME(da_t)* da = (NAME(da_t)*)malloc(sizeof(NAME(da_t)) * (size_t)C * (size_t)n_blocks);
  int max_depth = 1;
  for (int b = 0; b < n_blocks; ++b) {
    const REAL eps0 = blocks[b].kind == PM4_BLOCK_RWM ? (REAL)1 : (REAL)blocks[b].step_size;
    for (int c = 0; c < C; ++c) {
      NAME(da_init)(&da[(size_t)b * C + c], eps0);
      if (blocks[b].kind != PM4_BLOCK_RWM && blocks[b].adapt.shrinkage_target > 0)
        da[(size_t)b * C + c].log_shrink_target = (REAL)log(blocks[b].adapt.shrinkage_target);
    }
    if (blocks[b].kind == PM4_BLOCK_NUTS && blocks[b].max_tree_depth > max_depth) max_depth = blocks[b].max_tree_depth;
  }
  NAME(ws_t) ws = NAME(ws_new)(D);
  NAME(nuts_buf_t) nb = NAME(nuts_buf_new)(D, max_depth);
  REAL* base = nb.p0;
  for (int c = 0; c < C; ++c) lp[c] = NAME(logp_grad1)(ir, &ws, th + (size_t)c * D, gr + (size_t)c * D);
  for (int t = 0; t < T; ++t) {
    for (int b = 0; b < n_blocks; ++b) {
      const pm4_compound_block_t* k = &blocks[b];
      const uint64_t sb = seed + 0x9E3779B97F4A7C15ull * (uint64_t)b;
      const int adaptive = k->kind != PM4_BLOCK_RWM && k->adapt.enabled;
      const int pooled = adaptive && k->adapt.mode == PM4_EPS_POOLED;
      NAME(da_t)* dab = da + (size_t)b * C;
      pm4_nuts_cfg_t ncfg;
      memset(&ncfg, 0, sizeof ncfg);
      ncfg.C = C; ncfg.num_results = S; ncfg.num_burnin = B; ncfg.max_tree_depth = k->max_tree_depth;
      ncfg.max_energy_diff = k->max_energy_diff; ncfg.seed = sb; ncfg.adapt = k->adapt; ncfg.chain_offset = chain_offset;
      for (int c = 0; c < C; ++c) {
        const REAL eps = dab[pooled ? 0 : c].eps;
        REAL la;
        int nleap = 0;
        if (k->kind == PM4_BLOCK_NUTS) {
          NAME(nuts_out_t) o;
          NAME(nuts_transition)(ir, &ws, &nb, &ncfg, (uint32_t)(chain_offset + c), (uint32_t)t, eps,
                                (const REAL*)k->dim_scale_dev, NULL, th + (size_t)c * D, gr + (size_t)c * D, &lp[c], &o);
          la = o.log_accept;
          acc[c] = REXP(o.log_accept);
          nleap = o.leapfrogs;
        } else {
          int ok;
          NAME(mh_transition)(ir, &ws, k, sb, (uint32_t)(chain_offset + c), (uint32_t)t, eps, work, th + (size_t)c * D,
                              gr + (size_t)c * D, &lp[c], &la, &ok);
          acc[c] = REXP(la < 0 ? la : 0);
        }
        if (adaptive && !pooled) NAME(da_update)(&dab[c], &k->adapt, t, acc[c]);
        if (t >= B) {
          const size_t kk = (size_t)(t - B) * C + c;
          if (log_accept) log_accept[(size_t)b * S * C + kk] = la;
          if (leapfrogs_out) leapfrogs_out[(size_t)b * S * C + kk] = nleap;
          if (b == n_blocks - 1) {
            if (states) memcpy(states + kk * D, th + (size_t)c * D, sizeof(REAL) * (size_t)D);
            if (lp_out) lp_out[kk] = lp[c];
          }
        }
      }
      if (pooled && t <= k->adapt.num_adaptation_steps) { /* the device couples transitions 0 .. num_adapt */
        REAL s = 0;
        for (int c = 0; c < C; ++c) s += acc[c];
        NAME(da_update)(&dab[0], &k->adapt, t, s / (REAL)C);
      }
    }
  }
  if (step_size_out)
    for (int b = 0; b < n_blocks; ++b) {
      const int pooled = blocks[b].kind != PM4_BLOCK_RWM && blocks[b].adapt.enabled && blocks[b].adapt.mode == PM4_EPS_POOLED;
      for (int c = 0; c < C; ++c) step_size_out[(size_t)b * C + c] = da[(size_t)b * C + (pooled ? 0 : c)].eps;
    }
  nb.p0 = base;
  NAME(nuts_buf_free)(&nb);
  NAME(ws_free)(&ws);
  free(work); free(th); free(gr); free(lp); free(acc); free(da);
  return PM4_OK;
}
