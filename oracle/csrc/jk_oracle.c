/* CPU restatement (plain C + OpenMP) of the SQCC J/K contraction.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/code/python3/hf_ksdft.py:
 *   J[p,q] = sum_rs ERI[p,q,r,s] * D[r,s]      :483-487 (RHF/RKS), :510-515 (UHF/UKS, D = Da + Db)
 *   K[p,q] = sum_rs ERI[p,r,q,s] * D[r,s]      :491-495 (RHF), :522-526 (alpha), :531-535 (beta)
 * The loop nest (p, q outer; r, s inner, row-major) is the reference's; only the inner np.sum is a
 * plain long-double accumulation here (more accurate than numpy's pairwise double sum, so the oracle
 * error is negligible against the 1e-11 parity tolerance).
 *
 * Three operand sources so that config sizes fit in host memory:
 *   *_full     : the full [N,N,N,N] tensor, as interface_psi4.py:177-190 returns it;
 *   *_packed   : the canonical 8-fold packed vector (ij(ij+1)/2+kl), gathered per element;
 *   *_synth    : the counter-based synthetic tensor (oracle/synth.py), generated per element.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#include <stdint.h>
#include <stdlib.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int64_t pair_index(int64_t a, int64_t b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

static inline int64_t canon_offset(int64_t ab, int64_t cd)
{
    return ab >= cd ? ab * (ab + 1) / 2 + cd : cd * (cd + 1) / 2 + ab;
}

/* ---- synthetic generator: must match oracle/synth.py and csrc/eri_layout.cu bit for bit ---- */
static inline uint64_t splitmix64(uint64_t x)
{
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline double synth_value(int64_t n, uint64_t seed, int64_t i, int64_t j, int64_t k, int64_t l)
{
    /* i>=j, k>=l, ij>=kl */
    uint64_t np_ = (uint64_t)(n * (n + 1) / 2);
    uint64_t ij = (uint64_t)(i * (i + 1) / 2 + j), kl = (uint64_t)(k * (k + 1) / 2 + l);
    uint64_t z = splitmix64(seed + 0x9E3779B97F4A7C15ULL * (ij * np_ + kl + 1ULL));
    double u = (double)(z >> 11) * 0x1.0p-53;
    double t = 2.0 * u - 1.0;
    double d = (double)((i - j) + (k - l));
    double s = 1.0 / (1.0 + 0.25 * d);
    return ij == kl ? s * (fabs(t) + 1.0) : s * t;
}

static inline double synth_any(int64_t n, uint64_t seed, int64_t a, int64_t b, int64_t c, int64_t d)
{
    int64_t i = a >= b ? a : b, j = a >= b ? b : a, k = c >= d ? c : d, l = c >= d ? d : c;
    if (i * (i + 1) / 2 + j < k * (k + 1) / 2 + l) { int64_t t; t = i; i = k; k = t; t = j; j = l; l = t; }
    return synth_value(n, seed, i, j, k, l);
}

int sqcc_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void sqcc_oracle_synth_canonical(int64_t n, uint64_t seed, double *out)
{
    int64_t np_ = n * (n + 1) / 2;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j <= i; ++j) {
            int64_t ij = i * (i + 1) / 2 + j;
            double *row = out + ij * (ij + 1) / 2;
            for (int64_t k = 0; k <= i; ++k) {
                int64_t lmax = k < i ? k : j;
                for (int64_t l = 0; l <= lmax; ++l) row[k * (k + 1) / 2 + l] = synth_value(n, seed, i, j, k, l);
            }
        }
    (void)np_;
}


/* Rows [ij0, ij1) of the synthetic canonical packed tensor, stored back to back (row ij at out + off(ij) - off(ij0)). */
void sqcc_oracle_synth_rows(int64_t n, uint64_t seed, int64_t ij0, int64_t ij1, double *out)
{
    int64_t base = ij0 * (ij0 + 1) / 2;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t ij = ij0; ij < ij1; ++ij) {
        int64_t i = (int64_t)((sqrt(8.0 * (double)ij + 1.0) - 1.0) / 2.0);
        while (i * (i + 1) / 2 > ij) --i;
        while ((i + 1) * (i + 2) / 2 <= ij) ++i;
        int64_t j = ij - i * (i + 1) / 2;
        double *row = out + (ij * (ij + 1) / 2 - base);
        for (int64_t k = 0; k <= i; ++k) {
            int64_t lmax = k < i ? k : j;
            for (int64_t l = 0; l <= lmax; ++l) row[k * (k + 1) / 2 + l] = synth_value(n, seed, i, j, k, l);
        }
    }
}

/* source: 0 = full tensor, 1 = canonical packed, 2 = synthetic(seed) */
static inline double eri_at(int src, const double *buf, int64_t n, uint64_t seed, int64_t a, int64_t b, int64_t c,
                            int64_t d)
{
    if (src == 0) return buf[((a * n + b) * n + c) * n + d];
    if (src == 1) return buf[canon_offset(pair_index(a, b), pair_index(c, d))];
    return synth_any(n, seed, a, b, c, d);
}

/* J and K for the (p,q) entries listed in pq[2*cnt] (or all N*N when pq == NULL).
 * D is [nspin, N, N]; J out [cnt]; K out [nspin, cnt] (NULL to skip).  hf_ksdft.py:483-495 / :510-535. */
void sqcc_oracle_jk_entries(int src, const double *buf, int64_t n, uint64_t seed, const double *D, int nspin,
                            const int64_t *pq, int64_t cnt, double *J, double *K)
{
    double *dtot = (double *)malloc(sizeof(double) * n * n);
    for (int64_t t = 0; t < n * n; ++t) dtot[t] = nspin == 2 ? D[t] + D[n * n + t] : D[t]; /* :510 */
    if (pq == NULL) cnt = n * n;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t e = 0; e < cnt; ++e) {
        int64_t p = pq ? pq[2 * e] : e / n, q = pq ? pq[2 * e + 1] : e % n;
        long double sj = 0.0L, sk0 = 0.0L, sk1 = 0.0L;
        for (int64_t r = 0; r < n; ++r)
            for (int64_t s = 0; s < n; ++s) {
                sj += (long double)(eri_at(src, buf, n, seed, p, q, r, s) * dtot[r * n + s]);
                if (K) {
                    double x = eri_at(src, buf, n, seed, p, r, q, s);
                    sk0 += (long double)(x * D[r * n + s]);
                    if (nspin == 2) sk1 += (long double)(x * D[n * n + r * n + s]);
                }
            }
        J[e] = (double)sj;
        if (K) {
            K[e] = (double)sk0;
            if (nspin == 2) K[cnt + e] = (double)sk1;
        }
    }
    free(dtot);
}

/* Whole-matrix J/K from the canonical packed vector using the 8-fold accumulation of SURVEY.md A.1
 * (one pass over the packed vector, per-thread private J~/K~ then reduced).  This is the multi-core
 * "port" used as bench.py's cpu_baseline: same algorithm as the CUDA kernel, on the host cores. */
void sqcc_oracle_jk_packed_rows(const double *packed, int64_t n, int64_t ij0, int64_t ij1, const double *D, int nspin,
                                int want_k, double *J, double *K);

void sqcc_oracle_jk_packed(const double *packed, int64_t n, const double *D, int nspin, int want_k, double *J,
                           double *K)
{
    sqcc_oracle_jk_packed_rows(packed, n, 0, n * (n + 1) / 2, D, nspin, want_k, J, K);
}

/* Partial J/K from rows [ij0, ij1) only; `packed` points at the first element of row ij0. */
void sqcc_oracle_jk_packed_rows(const double *packed, int64_t n, int64_t ij0, int64_t ij1, const double *D, int nspin,
                                int want_k, double *J, double *K)
{
    int64_t nn = n * n;
    int nt = sqcc_oracle_num_threads();
    double *dtot = (double *)malloc(sizeof(double) * nn);
    for (int64_t t = 0; t < nn; ++t) dtot[t] = nspin == 2 ? D[t] + D[nn + t] : D[t];
    double *acc = (double *)calloc((size_t)nt * (size_t)(1 + nspin) * (size_t)nn, sizeof(double));
#pragma omp parallel
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        double *jt = acc + (size_t)tid * (1 + nspin) * nn;
        double *kt0 = jt + nn, *kt1 = jt + 2 * nn;
#pragma omp for schedule(dynamic, 8)
        for (int64_t ijr = 0; ijr < ij1 - ij0; ++ijr) {
            int64_t ij = ij1 - 1 - ijr; /* long rows first */
            int64_t i = (int64_t)((sqrt(8.0 * (double)ij + 1.0) - 1.0) / 2.0);
            while (i * (i + 1) / 2 > ij) --i;
            while ((i + 1) * (i + 2) / 2 <= ij) ++i;
            int64_t j = ij - i * (i + 1) / 2;
            const double *row = packed + (ij * (ij + 1) / 2 - ij0 * (ij0 + 1) / 2);
            double fij = i == j ? 0.5 : 1.0;
            double dij2 = dtot[i * n + j] + dtot[j * n + i];
            double jrow = 0.0;
            for (int64_t k = 0; k <= i; ++k) {
                int64_t lmax = k < i ? k : j;
                const double *sub = row + k * (k + 1) / 2;
                double a_ik0 = 0.0, a_jk0 = 0.0, a_ik1 = 0.0, a_jk1 = 0.0;
                for (int64_t l = 0; l <= lmax; ++l) {
                    double v = sub[l] * fij * (k == l ? 0.5 : 1.0) * ((k == i && l == j) ? 0.5 : 1.0);
                    jrow += v * (dtot[k * n + l] + dtot[l * n + k]);
                    jt[k * n + l] += v * dij2;
                    if (want_k) {
                        a_ik0 += v * D[j * n + l];
                        kt0[i * n + l] += v * D[j * n + k];
                        a_jk0 += v * D[i * n + l];
                        kt0[j * n + l] += v * D[i * n + k];
                        if (nspin == 2) {
                            a_ik1 += v * D[nn + j * n + l];
                            kt1[i * n + l] += v * D[nn + j * n + k];
                            a_jk1 += v * D[nn + i * n + l];
                            kt1[j * n + l] += v * D[nn + i * n + k];
                        }
                    }
                }
                if (want_k) {
                    kt0[i * n + k] += a_ik0;
                    kt0[j * n + k] += a_jk0;
                    if (nspin == 2) { kt1[i * n + k] += a_ik1; kt1[j * n + k] += a_jk1; }
                }
            }
            jt[i * n + j] += jrow;
        }
    }
    for (int64_t p = 0; p < n; ++p)
        for (int64_t q = 0; q < n; ++q) {
            long double sj = 0, s0 = 0, s1 = 0;
            for (int t = 0; t < nt; ++t) {
                const double *a = acc + (size_t)t * (1 + nspin) * nn;
                sj += a[p * n + q] + a[q * n + p];
                if (want_k) {
                    s0 += a[nn + p * n + q] + a[nn + q * n + p];
                    if (nspin == 2) s1 += a[2 * nn + p * n + q] + a[2 * nn + q * n + p];
                }
            }
            J[p * n + q] = (double)sj;
            if (want_k) {
                K[p * n + q] = (double)s0;
                if (nspin == 2) K[nn + p * n + q] = (double)s1;
            }
        }
    free(acc);
    free(dtot);
}

/* ---- MP2 AO->MO quarter transforms on the synthetic tensor (mp2.py:123-170), config-size oracle ----------------
 * (ia|jb) for ONE occupied pair (i, j) and a list of virtual pairs (a, b), evaluated with the reference's nesting of
 * the four np.dot steps restricted to the indices that are needed:
 *   temp1[i,q,r,s] = C[:,i] . ERI[:,q,r,s]      :123-130      temp2[i,q,j,s] = C[:,j] . temp1[i,q,:,s]   :136-143
 *   temp3[i,a,j,s] = C[:,a] . temp2[i,:,j,s]    :149-156      iajb[i,a,j,b]  = C[:,b] . temp3[i,a,j,:]   :162-170
 * C is the [N,N] AO x MO matrix (row-major); ci, cj, ca[], cb[] are MO column indices.  The tensor is generated
 * lazily (synth_any), so N = 222 needs no 19 GB host array.  Sums are long double (oracle accuracy). */
void sqcc_oracle_mp2_entries_synth(int64_t n, uint64_t seed, const double *C, int64_t ci, int64_t cj, const int64_t *ca,
                                   const int64_t *cb, int64_t cnt, double *out)
{
    double *t2 = (double *)malloc(sizeof(double) * n * n); /* temp2[i, q, j, s] for the fixed (i, j) */
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t qs = 0; qs < n * n; ++qs) {
        int64_t q = qs / n, s = qs % n;
        long double acc = 0.0L;
        for (int64_t r = 0; r < n; ++r) {
            long double t1 = 0.0L;
            for (int64_t p = 0; p < n; ++p) t1 += (long double)(C[p * n + ci] * synth_any(n, seed, p, q, r, s));
            acc += (long double)C[r * n + cj] * t1;
        }
        t2[qs] = (double)acc;
    }
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < cnt; ++e) {
        long double acc = 0.0L;
        for (int64_t s = 0; s < n; ++s) {
            long double t3 = 0.0L;
            for (int64_t q = 0; q < n; ++q) t3 += (long double)(C[q * n + ca[e]] * t2[q * n + s]);
            acc += (long double)C[s * n + cb[e]] * t3;
        }
        out[e] = (double)acc;
    }
    free(t2);
}

/* Slabs ERI[:, :, r, s] ([N, N] each, plain integral values) of the synthetic tensor for the canonical pairs
 * rs in [rs0, rs1): operands of the chunked einsum restatement of the quarter transforms (oracle/mp2_ref.py). */
void sqcc_oracle_synth_slabs_rs(int64_t n, uint64_t seed, int64_t rs0, int64_t rs1, double *out)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t rs = rs0; rs < rs1; ++rs) {
        int64_t r = (int64_t)((sqrt(8.0 * (double)rs + 1.0) - 1.0) / 2.0);
        while (r * (r + 1) / 2 > rs) --r;
        while ((r + 1) * (r + 2) / 2 <= rs) ++r;
        int64_t s = rs - r * (r + 1) / 2;
        double *slab = out + (rs - rs0) * n * n;
        for (int64_t p = 0; p < n; ++p)
            for (int64_t q = 0; q <= p; ++q) {
                double v = synth_any(n, seed, p, q, r, s);
                slab[p * n + q] = v;
                slab[q * n + p] = v;
            }
    }
}
