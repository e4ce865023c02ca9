/* ORACLE / CPU BASELINE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's CPU path for BLS12-381 G1, used (a) as a second, independent
 * checker next to oracle/bls12381.py and (b) as the `cpu_baseline` / `--impl reference` arm of bench.py,
 * timed on the GPU box's host cores.  Only tests/, __graft_entry__.smoke() and bench.py may load it.
 *
 * What the reference does on this path (it cannot be compiled here: no Go toolchain, and the field
 * code lives in the un-vendored module github.com/kilic/bls12-381 v0.1.0, go.mod:7):
 *   - Point.Mul: kilic.G1Elt.Mul -> G1.MulScalarBig      pairing/bls12381/kilic/g1.go:110-116
 *     one variable-time windowed double-and-add per call on 6x64-bit Montgomery limbs;
 *   - "MSM": there is none -- callers loop  Tmp.Mul(s_i, P_i); Acc.Add(Acc, Tmp)
 *     share/poly.go:461-473, sign/bdn/bdn.go:126-161.
 * Restated here with the same algorithm class: 6x64-bit CIOS Montgomery (unsigned __int128),
 * Jacobian coordinates, width-5 wNAF scalar multiplication (no GLV: stated, it is a <2x factor), and
 * the Mul+Add loop.  A multi-threaded Pippenger (cpu_g1_msm_pippenger) is provided as the STRONGER
 * CPU baseline the reference does not have.  PARITY: pinned against oracle/bls12381.py, which is
 * pinned against the reference's fixtures (tests/test_oracle_*.py); G1 bytes are canonical, so any
 * correct implementation is byte-identical to the Go path.
 *
 * Build: gcc -O3 -march=native -shared -fPIC -pthread oracle/cpu_ref.c -o oracle/libcpu_ref.so
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t fp[6];

static const fp P = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                     0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const uint64_t R_ORDER[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                                    0x73eda753299d7d48ULL};
static uint64_t M0;        /* -p^-1 mod 2^64 */
static fp R1, R2;          /* R mod p, R^2 mod p  (R = 2^384) */
static int g_init = 0;

static int fp_geq_p(const fp a) {
  for (int i = 5; i >= 0; i--) {
    if (a[i] > P[i]) return 1;
    if (a[i] < P[i]) return 0;
  }
  return 1;
}
static void fp_sub_p(fp a) {
  u128 br = 0;
  for (int i = 0; i < 6; i++) {
    u128 t = (u128)a[i] - P[i] - (uint64_t)br;
    a[i] = (uint64_t)t;
    br = (t >> 64) & 1;
  }
}
static void fp_add(fp r, const fp a, const fp b) {
  u128 c = 0;
  for (int i = 0; i < 6; i++) {
    c += (u128)a[i] + b[i];
    r[i] = (uint64_t)c;
    c >>= 64;
  }
  if (fp_geq_p(r)) fp_sub_p(r);
}
static void fp_sub(fp r, const fp a, const fp b) {
  u128 br = 0;
  for (int i = 0; i < 6; i++) {
    u128 t = (u128)a[i] - b[i] - (uint64_t)br;
    r[i] = (uint64_t)t;
    br = (t >> 64) & 1;
  }
  if (br) {
    u128 c = 0;
    for (int i = 0; i < 6; i++) {
      c += (u128)r[i] + P[i];
      r[i] = (uint64_t)c;
      c >>= 64;
    }
  }
}
static int fp_is_zero(const fp a) { return (a[0] | a[1] | a[2] | a[3] | a[4] | a[5]) == 0; }
static void fp_neg(fp r, const fp a) {
  if (fp_is_zero(a)) { memset(r, 0, sizeof(fp)); return; }
  fp z = {0};
  fp_sub(r, z, a);
}
/* CIOS Montgomery product, fully unrolled by the compiler */
#define UNROLL6 _Pragma("GCC unroll 6")
static inline __attribute__((always_inline)) void fp_mul(fp r, const fp a, const fp b) {
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0, t7;
  UNROLL6
  for (int i = 0; i < 6; i++) {
    const uint64_t bi = b[i];
    u128 c;
    c = (u128)a[0] * bi + t0; t0 = (uint64_t)c; c >>= 64;
    c += (u128)a[1] * bi + t1; t1 = (uint64_t)c; c >>= 64;
    c += (u128)a[2] * bi + t2; t2 = (uint64_t)c; c >>= 64;
    c += (u128)a[3] * bi + t3; t3 = (uint64_t)c; c >>= 64;
    c += (u128)a[4] * bi + t4; t4 = (uint64_t)c; c >>= 64;
    c += (u128)a[5] * bi + t5; t5 = (uint64_t)c; c >>= 64;
    c += t6; t6 = (uint64_t)c; t7 = (uint64_t)(c >> 64);
    const uint64_t m = t0 * M0;
    c = ((u128)m * P[0] + t0) >> 64;
    c += (u128)m * P[1] + t1; t0 = (uint64_t)c; c >>= 64;
    c += (u128)m * P[2] + t2; t1 = (uint64_t)c; c >>= 64;
    c += (u128)m * P[3] + t3; t2 = (uint64_t)c; c >>= 64;
    c += (u128)m * P[4] + t4; t3 = (uint64_t)c; c >>= 64;
    c += (u128)m * P[5] + t5; t4 = (uint64_t)c; c >>= 64;
    c += t6; t5 = (uint64_t)c; t6 = t7 + (uint64_t)(c >> 64);
  }
  r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3; r[4] = t4; r[5] = t5;
  if (t6 || fp_geq_p(r)) fp_sub_p(r);
}
static void fp_sqr(fp r, const fp a) { fp_mul(r, a, a); }
static void fp_pow(fp r, const fp a, const uint64_t* e, int nlimbs) {
  fp acc, base;
  memcpy(acc, R1, sizeof(fp));
  memcpy(base, a, sizeof(fp));
  for (int i = nlimbs * 64 - 1; i >= 0; i--) {
    fp_sqr(acc, acc);
    if ((e[i >> 6] >> (i & 63)) & 1) fp_mul(acc, acc, base);
  }
  memcpy(r, acc, sizeof(fp));
}
static void fp_inv(fp r, const fp a) {
  uint64_t e[6];
  memcpy(e, P, sizeof e);
  e[0] -= 2;
  fp_pow(r, a, e, 6);
}
static void fp_from_be(fp r, const uint8_t* b) {
  for (int i = 0; i < 6; i++) {
    uint64_t v = 0;
    for (int k = 0; k < 8; k++) v = (v << 8) | b[(5 - i) * 8 + k];
    r[i] = v;
  }
}
static void fp_to_be(uint8_t* b, const fp a) {
  for (int i = 0; i < 6; i++)
    for (int k = 0; k < 8; k++) b[(5 - i) * 8 + k] = (uint8_t)(a[i] >> (56 - 8 * k));
}

static void init_once(void) {
  if (g_init) return;
  /* M0 by Newton iteration */
  uint64_t inv = 1;
  for (int i = 0; i < 6; i++) inv *= 2 - P[0] * inv;
  M0 = (uint64_t)(0 - inv);
  /* R mod p by 384 modular doublings of 1; R^2 by 384 more */
  fp x = {1, 0, 0, 0, 0, 0};
  for (int i = 0; i < 768; i++) {
    fp_add(x, x, x);
    if (i == 383) memcpy(R1, x, sizeof(fp));
  }
  memcpy(R2, x, sizeof(fp));
  g_init = 1;
}

/* ---- G1, Jacobian; infinity Z = 0 ------------------------------------------------------------ */
typedef struct { fp X, Y, Z; } jac;
typedef struct { fp x, y; int inf; } aff;

static void jac_set_inf(jac* r) { memcpy(r->X, R1, sizeof(fp)); memcpy(r->Y, R1, sizeof(fp)); memset(r->Z, 0, sizeof(fp)); }
static void jac_dbl(jac* r, const jac* p) {
  if (fp_is_zero(p->Z)) { *r = *p; return; }
  fp A, B, C, D, E, F, T;
  fp_sqr(A, p->X); fp_sqr(B, p->Y); fp_sqr(C, B);
  fp_add(D, p->X, B); fp_sqr(D, D); fp_sub(D, D, A); fp_sub(D, D, C); fp_add(D, D, D);
  fp_add(E, A, A); fp_add(E, E, A);
  fp_sqr(F, E);
  fp_mul(T, p->Y, p->Z);
  fp_sub(F, F, D); fp_sub(F, F, D);
  fp_add(C, C, C); fp_add(C, C, C); fp_add(C, C, C);
  fp_sub(D, D, F); fp_mul(D, E, D); fp_sub(r->Y, D, C);
  memcpy(r->X, F, sizeof(fp));
  fp_add(r->Z, T, T);
}
static void jac_add(jac* r, const jac* p, const jac* q) {
  if (fp_is_zero(q->Z)) { *r = *p; return; }
  if (fp_is_zero(p->Z)) { *r = *q; return; }
  fp Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, T;
  fp_sqr(Z1Z1, p->Z); fp_sqr(Z2Z2, q->Z);
  fp_mul(U1, p->X, Z2Z2); fp_mul(U2, q->X, Z1Z1);
  fp_mul(S1, p->Y, q->Z); fp_mul(S1, S1, Z2Z2);
  fp_mul(S2, q->Y, p->Z); fp_mul(S2, S2, Z1Z1);
  fp_sub(H, U2, U1); fp_sub(rr, S2, S1);
  if (fp_is_zero(H)) {
    if (fp_is_zero(rr)) { jac_dbl(r, p); return; }
    jac_set_inf(r); return;
  }
  fp_add(rr, rr, rr);
  fp_add(I, H, H); fp_sqr(I, I);
  fp_mul(J, H, I); fp_mul(V, U1, I);
  fp_add(T, p->Z, q->Z); fp_sqr(T, T); fp_sub(T, T, Z1Z1); fp_sub(T, T, Z2Z2); fp_mul(T, T, H);
  fp_sqr(U2, rr); fp_sub(U2, U2, J); fp_sub(U2, U2, V); fp_sub(U2, U2, V);
  fp_sub(V, V, U2); fp_mul(V, rr, V);
  fp_mul(S1, S1, J); fp_add(S1, S1, S1);
  fp_sub(r->Y, V, S1);
  memcpy(r->X, U2, sizeof(fp));
  memcpy(r->Z, T, sizeof(fp));
}
static void jac_madd(jac* r, const jac* p, const aff* q, int negate) {
  if (q->inf) { *r = *p; return; }
  jac t;
  memcpy(t.X, q->x, sizeof(fp));
  if (negate) fp_neg(t.Y, q->y); else memcpy(t.Y, q->y, sizeof(fp));
  memcpy(t.Z, R1, sizeof(fp));
  jac_add(r, p, &t);
}
static void jac_to_aff(aff* r, const jac* p) {
  if (fp_is_zero(p->Z)) { memset(r, 0, sizeof *r); r->inf = 1; return; }
  fp zi, zi2;
  fp_inv(zi, p->Z); fp_sqr(zi2, zi);
  fp_mul(r->x, p->X, zi2); fp_mul(zi2, zi2, zi); fp_mul(r->y, p->Y, zi2);
  r->inf = 0;
}
static void aff_load(aff* r, const uint8_t* b /*96: x||y big-endian canonical, zero = inf*/) {
  int z = 1;
  for (int i = 0; i < 96; i++) if (b[i]) { z = 0; break; }
  if (z) { memset(r, 0, sizeof *r); r->inf = 1; return; }
  fp x, y;
  fp_from_be(x, b); fp_from_be(y, b + 48);
  fp_mul(r->x, x, R2); fp_mul(r->y, y, R2);
  r->inf = 0;
}
/* ZCash compressed 48 B (MarshalBinary, kilic/g1.go:119-124) */
static void aff_store_compressed(uint8_t* out, const aff* a) {
  if (a->inf) { memset(out, 0, 48); out[0] = 0xC0; return; }
  fp one = {1, 0, 0, 0, 0, 0}, x, y, ny;
  fp_mul(x, a->x, one); fp_mul(y, a->y, one);
  fp_to_be(out, x);
  /* y > (p-1)/2  <=>  y > p - y */
  fp z = {0};
  fp_sub(ny, z, y);
  int larger = 0;
  for (int i = 5; i >= 0; i--) { if (y[i] != ny[i]) { larger = y[i] > ny[i]; break; } }
  out[0] |= 0x80 | (larger ? 0x20 : 0);
}
/* operand form 96 B x||y (what the engine's *_affine entry points return); infinity = all zero */
static void aff_store_affine(uint8_t* out, const aff* a) {
  if (a->inf) { memset(out, 0, 96); return; }
  fp one = {1, 0, 0, 0, 0, 0}, x, y;
  fp_mul(x, a->x, one); fp_mul(y, a->y, one);
  fp_to_be(out, x); fp_to_be(out + 48, y);
}
static void scalar_from_be(uint64_t k[4], const uint8_t* b) {
  for (int i = 0; i < 4; i++) {
    uint64_t v = 0;
    for (int j = 0; j < 8; j++) v = (v << 8) | b[(3 - i) * 8 + j];
    k[i] = v;
  }
}
static int scalar_in_range(const uint64_t k[4]) {
  for (int i = 3; i >= 0; i--) {
    if (k[i] < R_ORDER[i]) return 1;
    if (k[i] > R_ORDER[i]) return 0;
  }
  return 0;
}

/* width-5 wNAF variable-time scalar multiplication (the algorithm class of the Go back-ends) */
static void g1_mul_wnaf(jac* r, const uint64_t k_in[4], const aff* p) {
  jac_set_inf(r);
  if (p->inf) return;
  int8_t naf[260];
  int len = 0;
  uint64_t k[5] = {k_in[0], k_in[1], k_in[2], k_in[3], 0};
  while (k[0] | k[1] | k[2] | k[3] | k[4]) {
    int d = 0;
    if (k[0] & 1) {
      d = (int)(k[0] & 31);
      if (d > 16) d -= 32;
      if (d > 0) { /* k -= d */
        u128 br = 0; uint64_t sub = (uint64_t)d;
        for (int i = 0; i < 5; i++) { u128 t = (u128)k[i] - sub - (uint64_t)br; k[i] = (uint64_t)t; br = (t >> 64) & 1; sub = 0; }
      } else {     /* k += -d */
        u128 c = (uint64_t)(-d);
        for (int i = 0; i < 5; i++) { c += k[i]; k[i] = (uint64_t)c; c >>= 64; }
      }
    }
    naf[len++] = (int8_t)d;
    for (int i = 0; i < 4; i++) k[i] = (k[i] >> 1) | (k[i + 1] << 63);
    k[4] >>= 1;
  }
  jac tbl[8], p2, pj;   /* 1P,3P,...,15P */
  memcpy(pj.X, p->x, sizeof(fp)); memcpy(pj.Y, p->y, sizeof(fp)); memcpy(pj.Z, R1, sizeof(fp));
  tbl[0] = pj;
  jac_dbl(&p2, &pj);
  for (int i = 1; i < 8; i++) jac_add(&tbl[i], &tbl[i - 1], &p2);
  for (int i = len - 1; i >= 0; i--) {
    jac_dbl(r, r);
    int d = naf[i];
    if (d > 0) jac_add(r, r, &tbl[d >> 1]);
    else if (d < 0) { jac t = tbl[(-d) >> 1]; fp_neg(t.Y, t.Y); jac_add(r, r, &t); }
  }
}

/* ---- single-threaded Pippenger over a range ---------------------------------------------------- */
static void pippenger_range(jac* out, size_t n, const uint8_t* scalars, const uint8_t* points) {
  jac_set_inf(out);
  if (n == 0) return;
  int c = 4;
  { double best = 1e300;
    for (int cc = 3; cc <= 16; cc++) { int W = (255 + cc) / cc; double cost = (double)W * ((double)n + 2.0 * (double)(1u << (cc - 1))); if (cost < best) { best = cost; c = cc; } } }
  int W = (255 + c) / c;   /* covers 255 bits + carry */
  size_t nb = (size_t)1 << (c - 1);
  aff* pts = (aff*)malloc(n * sizeof(aff));
  int32_t* dig = (int32_t*)malloc(n * (size_t)W * sizeof(int32_t));
  for (size_t i = 0; i < n; i++) {
    aff_load(&pts[i], points + 96 * i);
    uint64_t k[4];
    scalar_from_be(k, scalars + 32 * i);
    int carry = 0;
    for (int w = 0; w < W; w++) {
      int bit = w * c;
      uint64_t v = 0;
      if (bit < 256) {
        v = k[bit >> 6] >> (bit & 63);
        if ((bit & 63) + c > 64 && (bit >> 6) + 1 < 4) v |= k[(bit >> 6) + 1] << (64 - (bit & 63));
        v &= ((uint64_t)1 << c) - 1;
      }
      int d = (int)v + carry;
      carry = 0;
      if (d > (int)nb) { d -= (1 << c); carry = 1; }
      dig[i * W + w] = d;
    }
  }
  jac* buckets = (jac*)malloc(nb * sizeof(jac));
  jac total;
  jac_set_inf(&total);
  for (int w = W - 1; w >= 0; w--) {
    for (int i = 0; i < c; i++) jac_dbl(&total, &total);
    for (size_t b = 0; b < nb; b++) jac_set_inf(&buckets[b]);
    for (size_t i = 0; i < n; i++) {
      int d = dig[i * W + w];
      if (d > 0) jac_madd(&buckets[d - 1], &buckets[d - 1], &pts[i], 0);
      else if (d < 0) jac_madd(&buckets[-d - 1], &buckets[-d - 1], &pts[i], 1);
    }
    jac run, acc;
    jac_set_inf(&run); jac_set_inf(&acc);
    for (size_t b = nb; b-- > 0;) { jac_add(&run, &run, &buckets[b]); jac_add(&acc, &acc, &run); }
    jac_add(&total, &total, &acc);
  }
  *out = total;
  free(buckets); free(dig); free(pts);
}

/* ---- threading ------------------------------------------------------------------------------------ */
typedef struct {
  int mode;  /* 0 = mul batch, 1 = Mul+Add loop partial, 2 = pippenger partial, 3 = mul batch with operand-form output */
  size_t lo, hi;
  const uint8_t *scalars, *points;
  uint8_t* out;
  jac partial;
  int bad;
} job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  j->bad = 0;
  if (j->mode == 2) {
    for (size_t i = j->lo; i < j->hi; i++) { uint64_t k[4]; scalar_from_be(k, j->scalars + 32 * i); if (!scalar_in_range(k)) j->bad = 1; }
    pippenger_range(&j->partial, j->hi - j->lo, j->scalars + 32 * j->lo, j->points + 96 * j->lo);
    return NULL;
  }
  jac_set_inf(&j->partial);
  for (size_t i = j->lo; i < j->hi; i++) {
    uint64_t k[4];
    scalar_from_be(k, j->scalars + 32 * i);
    if (!scalar_in_range(k)) j->bad = 1;
    aff p;
    aff_load(&p, j->points + 96 * i);
    jac r;
    g1_mul_wnaf(&r, k, &p);
    if (j->mode == 0) { aff a; jac_to_aff(&a, &r); aff_store_compressed(j->out + 48 * i, &a); }
    else if (j->mode == 3) { aff a; jac_to_aff(&a, &r); aff_store_affine(j->out + 96 * i, &a); }
    else jac_add(&j->partial, &j->partial, &r);   /* Acc.Add(Acc, Tmp)  share/poly.go:472 */
  }
  return NULL;
}

static int run_jobs(int mode, size_t n, const uint8_t* scalars, const uint8_t* points, uint8_t* out, int nthreads, jac* sum) {
  init_once();
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n) nthreads = (int)n;
  job_t* jobs = (job_t*)calloc((size_t)nthreads, sizeof(job_t));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; t++) {
    jobs[t].mode = mode; jobs[t].lo = n * (size_t)t / nthreads; jobs[t].hi = n * (size_t)(t + 1) / nthreads;
    jobs[t].scalars = scalars; jobs[t].points = points; jobs[t].out = out;
    if (nthreads == 1) worker(&jobs[t]); else pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  int bad = 0;
  if (sum) jac_set_inf(sum);
  for (int t = 0; t < nthreads; t++) {
    if (nthreads > 1) pthread_join(th[t], NULL);
    bad |= jobs[t].bad;
    if (sum) jac_add(sum, sum, &jobs[t].partial);
  }
  free(jobs); free(th);
  return bad ? -3 : 0;
}

/* n independent Point.Mul -> 48-byte MarshalBinary each */
int cpu_g1_mul_batch(size_t n, const uint8_t* scalars, const uint8_t* points, uint8_t* out48n, int nthreads) {
  return run_jobs(0, n, scalars, points, out48n, nthreads, NULL);
}
/* the same with 96-byte operand-form results (bench.py builds its CPU sample with it: no per-point Python) */
int cpu_g1_mul_batch_affine(size_t n, const uint8_t* scalars, const uint8_t* points, uint8_t* out96n, int nthreads) {
  return run_jobs(3, n, scalars, points, out96n, nthreads, NULL);
}
/* the reference's way to do an MSM: loop of Mul + Add (share/poly.go:461-473) */
int cpu_g1_msm_muladd(size_t n, const uint8_t* scalars, const uint8_t* points, uint8_t* out48, int nthreads) {
  jac s; aff a;
  int rc = run_jobs(1, n, scalars, points, NULL, nthreads, &s);
  jac_to_aff(&a, &s); aff_store_compressed(out48, &a);
  return rc;
}
/* stronger CPU baseline: Pippenger, point range split across threads */
int cpu_g1_msm_pippenger(size_t n, const uint8_t* scalars, const uint8_t* points, uint8_t* out48, int nthreads) {
  jac s; aff a;
  int rc = run_jobs(2, n, scalars, points, NULL, nthreads, &s);
  jac_to_aff(&a, &s); aff_store_compressed(out48, &a);
  return rc;
}

/* ==================================================================================================
 * Pairing (CPU baseline for pairings/s).  Restates the optimal-ate pairing the reference reaches through
 * kilic.Suite.Pair / ValidatePairing (pairing/bls12381/kilic/suite.go:57-75; arithmetic third-party),
 * structured like the in-tree bn254 one (pairing/bn254/optate.go:5-271): Jacobian twist point,
 * inversion-free lines, sparse line multiplication, final exponentiation with exponent (p^12-1)/r.
 * Checked against oracle/bls12381.py (GT bytes) in tests/test_oracle_bls12381.py.
 * ================================================================================================== */
typedef struct { fp c0, c1; } fp2;
typedef struct { fp2 c0, c1, c2; } fp6;
typedef struct { fp6 c0, c1; } fp12;

static void fp2_add(fp2* r, const fp2* a, const fp2* b) { fp_add(r->c0, a->c0, b->c0); fp_add(r->c1, a->c1, b->c1); }
static void fp2_sub(fp2* r, const fp2* a, const fp2* b) { fp_sub(r->c0, a->c0, b->c0); fp_sub(r->c1, a->c1, b->c1); }
static void fp2_neg(fp2* r, const fp2* a) { fp_neg(r->c0, a->c0); fp_neg(r->c1, a->c1); }
static void fp2_conj(fp2* r, const fp2* a) { memcpy(r->c0, a->c0, sizeof(fp)); fp_neg(r->c1, a->c1); }
static void fp2_dbl(fp2* r, const fp2* a) { fp2_add(r, a, a); }
static void fp2_mul(fp2* r, const fp2* a, const fp2* b) {
  fp t0, t1, s0, s1;
  fp_mul(t0, a->c0, b->c0); fp_mul(t1, a->c1, b->c1);
  fp_add(s0, a->c0, a->c1); fp_add(s1, b->c0, b->c1); fp_mul(s0, s0, s1);
  fp_sub(s0, s0, t0); fp_sub(r->c1, s0, t1); fp_sub(r->c0, t0, t1);
}
static void fp2_sqr(fp2* r, const fp2* a) {
  fp s, d, m;
  fp_add(s, a->c0, a->c1); fp_sub(d, a->c0, a->c1); fp_mul(m, a->c0, a->c1);
  fp_mul(r->c0, s, d); fp_add(r->c1, m, m);
}
static void fp2_mul_fp(fp2* r, const fp2* a, const fp k) { fp_mul(r->c0, a->c0, k); fp_mul(r->c1, a->c1, k); }
static void fp2_mul_xi(fp2* r, const fp2* a) { fp t; fp_sub(t, a->c0, a->c1); fp_add(r->c1, a->c0, a->c1); memcpy(r->c0, t, sizeof(fp)); }
static void fp2_inv(fp2* r, const fp2* a) {
  fp n, t;
  fp_sqr(n, a->c0); fp_sqr(t, a->c1); fp_add(n, n, t); fp_inv(n, n);
  fp_mul(r->c0, a->c0, n); fp_mul(t, a->c1, n); fp_neg(r->c1, t);
}
static void fp6_add(fp6* r, const fp6* a, const fp6* b) { fp2_add(&r->c0, &a->c0, &b->c0); fp2_add(&r->c1, &a->c1, &b->c1); fp2_add(&r->c2, &a->c2, &b->c2); }
static void fp6_sub(fp6* r, const fp6* a, const fp6* b) { fp2_sub(&r->c0, &a->c0, &b->c0); fp2_sub(&r->c1, &a->c1, &b->c1); fp2_sub(&r->c2, &a->c2, &b->c2); }
static void fp6_neg(fp6* r, const fp6* a) { fp2_neg(&r->c0, &a->c0); fp2_neg(&r->c1, &a->c1); fp2_neg(&r->c2, &a->c2); }
static void fp6_mul_v(fp6* r, const fp6* a) { fp2 t; fp2_mul_xi(&t, &a->c2); r->c2 = a->c1; r->c1 = a->c0; r->c0 = t; }
static void fp6_mul(fp6* r, const fp6* a, const fp6* b) {
  fp2 v0, v1, v2, t0, t1, t2, s;
  fp2_mul(&v0, &a->c0, &b->c0); fp2_mul(&v1, &a->c1, &b->c1); fp2_mul(&v2, &a->c2, &b->c2);
  fp2_add(&t0, &a->c1, &a->c2); fp2_add(&s, &b->c1, &b->c2); fp2_mul(&t0, &t0, &s);
  fp2_sub(&t0, &t0, &v1); fp2_sub(&t0, &t0, &v2); fp2_mul_xi(&t0, &t0); fp2_add(&t0, &t0, &v0);
  fp2_add(&t1, &a->c0, &a->c1); fp2_add(&s, &b->c0, &b->c1); fp2_mul(&t1, &t1, &s);
  fp2_sub(&t1, &t1, &v0); fp2_sub(&t1, &t1, &v1); fp2_mul_xi(&s, &v2); fp2_add(&t1, &t1, &s);
  fp2_add(&t2, &a->c0, &a->c2); fp2_add(&s, &b->c0, &b->c2); fp2_mul(&t2, &t2, &s);
  fp2_sub(&t2, &t2, &v0); fp2_sub(&t2, &t2, &v2); fp2_add(&t2, &t2, &v1);
  r->c0 = t0; r->c1 = t1; r->c2 = t2;
}
static void fp6_inv(fp6* r, const fp6* a) {
  fp2 t0, t1, t2, s, d;
  fp2_sqr(&t0, &a->c0); fp2_mul(&s, &a->c1, &a->c2); fp2_mul_xi(&s, &s); fp2_sub(&t0, &t0, &s);
  fp2_sqr(&t1, &a->c2); fp2_mul_xi(&t1, &t1); fp2_mul(&s, &a->c0, &a->c1); fp2_sub(&t1, &t1, &s);
  fp2_sqr(&t2, &a->c1); fp2_mul(&s, &a->c0, &a->c2); fp2_sub(&t2, &t2, &s);
  fp2_mul(&d, &a->c2, &t1); fp2_mul(&s, &a->c1, &t2); fp2_add(&d, &d, &s); fp2_mul_xi(&d, &d);
  fp2_mul(&s, &a->c0, &t0); fp2_add(&d, &d, &s); fp2_inv(&d, &d);
  fp2_mul(&r->c0, &t0, &d); fp2_mul(&r->c1, &t1, &d); fp2_mul(&r->c2, &t2, &d);
}
static void fp12_one(fp12* r) { memset(r, 0, sizeof *r); memcpy(r->c0.c0.c0, R1, sizeof(fp)); }
static void fp12_mul(fp12* r, const fp12* a, const fp12* b) {
  fp6 t0, t1, s0, s1;
  fp6_mul(&t0, &a->c0, &b->c0); fp6_mul(&t1, &a->c1, &b->c1);
  fp6_add(&s0, &a->c0, &a->c1); fp6_add(&s1, &b->c0, &b->c1); fp6_mul(&s0, &s0, &s1);
  fp6_sub(&s0, &s0, &t0); fp6_sub(&r->c1, &s0, &t1);
  fp6_mul_v(&t1, &t1); fp6_add(&r->c0, &t0, &t1);
}
static void fp12_sqr(fp12* r, const fp12* a) {
  fp6 ab, s, t;
  fp6_mul(&ab, &a->c0, &a->c1);
  fp6_add(&s, &a->c0, &a->c1); fp6_mul_v(&t, &a->c1); fp6_add(&t, &t, &a->c0);
  fp6_mul(&s, &s, &t); fp6_sub(&s, &s, &ab); fp6_mul_v(&t, &ab);
  fp6_sub(&r->c0, &s, &t); fp6_add(&r->c1, &ab, &ab);
}
static void fp12_conj(fp12* r, const fp12* a) { r->c0 = a->c0; fp6_neg(&r->c1, &a->c1); }
static void fp12_inv(fp12* r, const fp12* a) {
  fp6 d, t;
  fp6_mul(&d, &a->c0, &a->c0); fp6_mul(&t, &a->c1, &a->c1); fp6_mul_v(&t, &t); fp6_sub(&d, &d, &t);
  fp6_inv(&d, &d);
  fp6_mul(&r->c0, &a->c0, &d); fp6_mul(&t, &a->c1, &d); fp6_neg(&r->c1, &t);
}
/* sparse line (l0, l2, 0) + (0, l3, 0) w */
static void fp12_mul_line(fp12* f, const fp2* l0, const fp2* l2, const fp2* l3) {
  fp12 l;
  memset(&l, 0, sizeof l);
  l.c0.c0 = *l0; l.c0.c1 = *l2; l.c1.c1 = *l3;
  fp12_mul(f, f, &l);      /* dense product: the Go back-ends use sparse formulas; <15% of a pairing */
}
static fp2 FROB1[6];
static int g_pinit = 0;
static void pairing_init(void) {
  if (g_pinit) return;
  init_once();
  /* (p-1)/6 by long division */
  uint64_t e[6]; u128 rem = 0;
  uint64_t pm1[6]; memcpy(pm1, P, sizeof pm1); pm1[0] -= 1;
  for (int i = 5; i >= 0; i--) { u128 cur = (rem << 64) | pm1[i]; e[i] = (uint64_t)(cur / 6); rem = cur % 6; }
  fp2 xi, acc, base;
  memcpy(xi.c0, R1, sizeof(fp)); memcpy(xi.c1, R1, sizeof(fp));
  memset(&acc, 0, sizeof acc); memcpy(acc.c0, R1, sizeof(fp));
  base = xi;
  for (int i = 6 * 64 - 1; i >= 0; i--) { fp2_sqr(&acc, &acc); if ((e[i >> 6] >> (i & 63)) & 1) fp2_mul(&acc, &acc, &base); }
  memset(&FROB1[0], 0, sizeof(fp2)); memcpy(FROB1[0].c0, R1, sizeof(fp));
  for (int k = 1; k < 6; k++) fp2_mul(&FROB1[k], &FROB1[k - 1], &acc);
  g_pinit = 1;
}
static void fp12_frob(fp12* r, const fp12* f) {   /* f^p */
  fp2* dst[6] = {&r->c0.c0, &r->c1.c0, &r->c0.c1, &r->c1.c1, &r->c0.c2, &r->c1.c2};
  const fp2* src[6] = {&f->c0.c0, &f->c1.c0, &f->c0.c1, &f->c1.c1, &f->c0.c2, &f->c1.c2};
  for (int k = 0; k < 6; k++) { fp2 a; fp2_conj(&a, src[k]); fp2_mul(dst[k], &a, &FROB1[k]); }
}
static void fp12_pow_u64(fp12* r, const fp12* a, uint64_t e) {
  fp12 acc = *a;
  int top = 63;
  while (top > 0 && !((e >> top) & 1)) top--;
  for (int b = top - 1; b >= 0; b--) { fp12_sqr(&acc, &acc); if ((e >> b) & 1) fp12_mul(&acc, &acc, a); }
  *r = acc;
}
#define X_ABS 0xd201000000010000ULL
static void fp12_pow_x(fp12* r, const fp12* a) { fp12_pow_u64(r, a, X_ABS); fp12_conj(r, r); }   /* a^x, x < 0 */
/* exponent 3 (p^12-1)/r, the GT convention of the reference back-ends (pinned by encrypt/ibe/ibe_test.go:202-245,
 * see oracle/bls12381.py pairing_reference): hard part (x-1)^2 (x+p) (x^2+p^2-1) + 3 */
static void final_exp(fp12* r, const fp12* f) {
  fp12 m, t, a, b, c;
  fp12_inv(&t, f); fp12_conj(&m, f); fp12_mul(&m, &m, &t);
  fp12_frob(&t, &m); fp12_frob(&t, &t); fp12_mul(&m, &t, &m);
  fp12_pow_x(&b, &m); fp12_conj(&t, &m); fp12_mul(&b, &b, &t);
  fp12_pow_x(&a, &b); fp12_conj(&t, &b); fp12_mul(&a, &a, &t);
  fp12_pow_x(&c, &a); fp12_frob(&t, &a); fp12_mul(&c, &c, &t);
  fp12_pow_x(&b, &c); fp12_pow_x(&a, &b);
  fp12_frob(&t, &c); fp12_frob(&t, &t); fp12_mul(&a, &a, &t);
  fp12_conj(&t, &c); fp12_mul(&a, &a, &t);
  fp12_sqr(&t, &m); fp12_mul(&t, &t, &m);
  fp12_mul(r, &a, &t);
}
typedef struct { fp2 X, Y, Z; } jac2;
typedef struct { fp2 x, y; int inf; } aff2;
static void g2_load(aff2* r, const uint8_t* b /*192: x.c1||x.c0||y.c1||y.c0*/) {
  int z = 1;
  for (int i = 0; i < 192; i++) if (b[i]) { z = 0; break; }
  memset(r, 0, sizeof *r);
  if (z) { r->inf = 1; return; }
  fp t;
  fp_from_be(t, b); fp_mul(r->x.c1, t, R2); fp_from_be(t, b + 48); fp_mul(r->x.c0, t, R2);
  fp_from_be(t, b + 96); fp_mul(r->y.c1, t, R2); fp_from_be(t, b + 144); fp_mul(r->y.c0, t, R2);
}
static void miller_dbl(fp2* l0, fp2* l2, fp2* l3, jac2* T, const aff* Pp) {
  fp2 A, B, C, D, E, ZZ, t;
  fp2_sqr(&A, &T->X); fp2_sqr(&B, &T->Y); fp2_sqr(&C, &B); fp2_sqr(&ZZ, &T->Z);
  fp2_add(&D, &T->X, &B); fp2_sqr(&D, &D); fp2_sub(&D, &D, &A); fp2_sub(&D, &D, &C); fp2_dbl(&D, &D);
  fp2_dbl(&E, &A); fp2_add(&E, &E, &A);
  fp2_mul(l0, &E, &T->X); fp2_sub(l0, l0, &B); fp2_sub(l0, l0, &B);
  fp2_mul(&t, &E, &ZZ); fp2_mul_fp(&t, &t, Pp->x); fp2_neg(l2, &t);
  fp2_mul(&t, &T->Y, &T->Z); fp2_dbl(&T->Z, &t);
  fp2_mul(&t, &T->Z, &ZZ); fp2_mul_fp(l3, &t, Pp->y);
  fp2_sqr(&A, &E); fp2_sub(&A, &A, &D); fp2_sub(&A, &A, &D);
  fp2_dbl(&C, &C); fp2_dbl(&C, &C); fp2_dbl(&C, &C);
  fp2_sub(&D, &D, &A); fp2_mul(&D, &E, &D); fp2_sub(&T->Y, &D, &C);
  T->X = A;
}
static void miller_add(fp2* l0, fp2* l2, fp2* l3, jac2* T, const aff2* Q, const aff* Pp) {
  fp2 ZZ, U2, S2, H, R, HH, HHH, V, t;
  fp2_sqr(&ZZ, &T->Z); fp2_mul(&U2, &Q->x, &ZZ); fp2_mul(&S2, &Q->y, &T->Z); fp2_mul(&S2, &S2, &ZZ);
  fp2_sub(&H, &U2, &T->X); fp2_sub(&R, &S2, &T->Y);
  fp2_sqr(&HH, &H); fp2_mul(&HHH, &H, &HH); fp2_mul(&V, &T->X, &HH);
  fp2_mul(&T->Z, &T->Z, &H);
  fp2_mul(l0, &R, &Q->x); fp2_mul(&t, &T->Z, &Q->y); fp2_sub(l0, l0, &t);
  fp2_mul_fp(&t, &R, Pp->x); fp2_neg(l2, &t);
  fp2_mul_fp(l3, &T->Z, Pp->y);
  fp2_sqr(&t, &R); fp2_sub(&t, &t, &HHH); fp2_sub(&t, &t, &V); fp2_sub(&t, &t, &V);
  fp2_sub(&V, &V, &t); fp2_mul(&V, &R, &V); fp2_mul(&HHH, &T->Y, &HHH); fp2_sub(&T->Y, &V, &HHH);
  T->X = t;
}
static void miller_n(fp12* f, int np, const aff* Pp, const aff2* Q) {
  jac2 T[2]; int live[2];
  for (int i = 0; i < np; i++) { live[i] = !(Pp[i].inf || Q[i].inf); T[i].X = Q[i].x; T[i].Y = Q[i].y; memset(&T[i].Z, 0, sizeof(fp2)); memcpy(T[i].Z.c0, R1, sizeof(fp)); }
  fp12_one(f);
  fp2 l0, l2, l3;
  for (int b = 62; b >= 0; b--) {
    fp12_sqr(f, f);
    for (int i = 0; i < np; i++) if (live[i]) { miller_dbl(&l0, &l2, &l3, &T[i], &Pp[i]); fp12_mul_line(f, &l0, &l2, &l3); }
    if ((X_ABS >> b) & 1) for (int i = 0; i < np; i++) if (live[i]) { miller_add(&l0, &l2, &l3, &T[i], &Q[i], &Pp[i]); fp12_mul_line(f, &l0, &l2, &l3); }
  }
  fp12_conj(f, f);
}
static void gt_store(uint8_t* out, const fp12* f) {
  const fp2* order[6] = {&f->c1.c2, &f->c1.c1, &f->c1.c0, &f->c0.c2, &f->c0.c1, &f->c0.c0};
  fp one = {1, 0, 0, 0, 0, 0}, t;
  for (int i = 0; i < 6; i++) { fp_mul(t, order[i]->c1, one); fp_to_be(out + 96 * i, t); fp_mul(t, order[i]->c0, one); fp_to_be(out + 96 * i + 48, t); }
}
typedef struct { size_t lo, hi; const uint8_t *a1, *a2, *b1, *b2; uint8_t* out; int mode; } pjob_t;
static void* pworker(void* arg) {
  pjob_t* j = (pjob_t*)arg;
  for (size_t i = j->lo; i < j->hi; i++) {
    aff Pp[2]; aff2 Q[2]; fp12 f, e;
    aff_load(&Pp[0], j->a1 + 96 * i); g2_load(&Q[0], j->a2 + 192 * i);
    if (j->mode == 0) { miller_n(&f, 1, Pp, Q); final_exp(&e, &f); gt_store(j->out + 576 * i, &e); }
    else {
      aff_load(&Pp[1], j->b1 + 96 * i); g2_load(&Q[1], j->b2 + 192 * i); fp_neg(Pp[1].y, Pp[1].y);
      miller_n(&f, 2, Pp, Q); final_exp(&e, &f);
      fp12 one; fp12_one(&one);
      j->out[i] = memcmp(&e, &one, sizeof one) == 0;
    }
  }
  return NULL;
}
static int prun(int mode, size_t n, const uint8_t* a1, const uint8_t* a2, const uint8_t* b1, const uint8_t* b2, uint8_t* out, int nthreads) {
  pairing_init();
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n) nthreads = (int)n;
  pjob_t* jobs = (pjob_t*)calloc((size_t)nthreads, sizeof(pjob_t));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (pjob_t){n * (size_t)t / nthreads, n * (size_t)(t + 1) / nthreads, a1, a2, b1, b2, out, mode};
    if (nthreads == 1) pworker(&jobs[t]); else pthread_create(&th[t], NULL, pworker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) if (nthreads > 1) pthread_join(th[t], NULL);
  free(jobs); free(th);
  return 0;
}
/* gt[i] = e(g1[i], g2[i]) */
int cpu_pair(size_t n, const uint8_t* g1, const uint8_t* g2, uint8_t* gt576n, int nthreads) { return prun(0, n, g1, g2, NULL, NULL, gt576n, nthreads); }
/* ok[i] = e(a1,a2) == e(b1,b2) */
int cpu_pairing_check(size_t n, const uint8_t* a1, const uint8_t* a2, const uint8_t* b1, const uint8_t* b2, uint8_t* ok, int nthreads) { return prun(1, n, a1, a2, b1, b2, ok, nthreads); }
