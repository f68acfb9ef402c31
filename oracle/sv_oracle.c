/*
 * sv_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see sv_oracle.h header comment).
 *
 * Gate-by-gate, unfused, fp64.  Every routine is the textbook index-insertion loop that the
 * reference's CPU simulators run behind the call sites cited per function below
 * (file = /root/reference/src/backends/simulators/AER/aer_adapters/aer_simulator_adapter.cpp).
 * PARITY UNPINNED (no golden vectors exist in the reference for this path).
 */
#include "sv_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXK 10

/* pin the OpenMP team size (launchers such as torch.distributed.run export OMP_NUM_THREADS=1: a CPU baseline timed
 * under them must ask for the host's cores itself); n <= 0: every online core.  Returns the team size now in force. */
int orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n <= 0) n = omp_get_num_procs();
    omp_set_dynamic(0);
    omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_init_zero(double *state, int n)
{
    const uint64_t dim = 1ull << n;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < dim; i++) { state[2 * i] = 0.0; state[2 * i + 1] = 0.0; }
    state[0] = 1.0;
}

static void sort_ints(int *a, int m)
{
    for (int i = 1; i < m; i++) {
        int v = a[i], j = i - 1;
        while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; j--; }
        a[j + 1] = v;
    }
}

/* insert a zero bit at every (ascending) position in pos[] */
static inline uint64_t insert_zeros(uint64_t g, const int *pos, int m)
{
    for (int i = 0; i < m; i++) {
        const uint64_t lo = g & ((1ull << pos[i]) - 1);
        g = ((g >> pos[i]) << (pos[i] + 1)) | lo;
    }
    return g;
}

/* follows apply_unitary / apply_mcu / apply_mcx... call sites (:203-470): new = M * old on the
 * 2^k sub-vector addressed by qubits[], restricted to the control=1 subspace. */
void orc_apply_matrix(double *state, int n, const int *qubits, int k,
                      const int *ctrls, int nc, const double *mat)
{
    if (k > MAXK) abort();
    const int dim = 1 << k;
    int pos[64];
    uint64_t cmask = 0;
    for (int i = 0; i < k; i++) pos[i] = qubits[i];
    for (int i = 0; i < nc; i++) { pos[k + i] = ctrls[i]; cmask |= 1ull << ctrls[i]; }
    sort_ints(pos, k + nc);
    uint64_t off[1 << MAXK];
    for (int j = 0; j < dim; j++) {
        uint64_t o = 0;
        for (int m = 0; m < k; m++) if ((j >> m) & 1) o |= 1ull << qubits[m];
        off[j] = o;
    }
    const uint64_t ngroups = 1ull << (n - k - nc);
#pragma omp parallel
    {
        double *in = (double *)malloc(sizeof(double) * 2 * dim);
#pragma omp for schedule(static)
        for (uint64_t g = 0; g < ngroups; g++) {
            const uint64_t base = insert_zeros(g, pos, k + nc) | cmask;
            for (int j = 0; j < dim; j++) {
                in[2 * j] = state[2 * (base | off[j])];
                in[2 * j + 1] = state[2 * (base | off[j]) + 1];
            }
            for (int r = 0; r < dim; r++) {
                double sr = 0.0, si = 0.0;
                const double *row = mat + 2 * (size_t)r * dim;
                for (int c = 0; c < dim; c++) {
                    const double mr = row[2 * c], mi = row[2 * c + 1];
                    sr += mr * in[2 * c] - mi * in[2 * c + 1];
                    si += mr * in[2 * c + 1] + mi * in[2 * c];
                }
                state[2 * (base | off[r])] = sr;
                state[2 * (base | off[r]) + 1] = si;
            }
        }
        free(in);
    }
}

/* follows apply_diagonal_matrix (:531) / apply_mcphase (:337) / apply_mcrz (:394) */
void orc_apply_diagonal(double *state, int n, const int *qubits, int k, const double *diag)
{
    const uint64_t dim = 1ull << n;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < dim; i++) {
        int j = 0;
        for (int m = 0; m < k; m++) j |= (int)((i >> qubits[m]) & 1) << m;
        const double dr = diag[2 * j], di = diag[2 * j + 1];
        const double ar = state[2 * i], ai = state[2 * i + 1];
        state[2 * i] = dr * ar - di * ai;
        state[2 * i + 1] = dr * ai + di * ar;
    }
}

/* follows apply_mcswap (:242,:432,:661) */
void orc_apply_swap(double *state, int n, int q0, int q1, const int *ctrls, int nc)
{
    int pos[64];
    uint64_t cmask = 0;
    pos[0] = q0; pos[1] = q1;
    for (int i = 0; i < nc; i++) { pos[2 + i] = ctrls[i]; cmask |= 1ull << ctrls[i]; }
    sort_ints(pos, 2 + nc);
    const uint64_t ngroups = 1ull << (n - 2 - nc);
#pragma omp parallel for schedule(static)
    for (uint64_t g = 0; g < ngroups; g++) {
        const uint64_t base = insert_zeros(g, pos, 2 + nc) | cmask;
        const uint64_t a = base | (1ull << q0), b = base | (1ull << q1);
        const double tr = state[2 * a], ti = state[2 * a + 1];
        state[2 * a] = state[2 * b]; state[2 * a + 1] = state[2 * b + 1];
        state[2 * b] = tr; state[2 * b + 1] = ti;
    }
}

double orc_norm(const double *state, int n)
{
    const uint64_t dim = 1ull << n;
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (uint64_t i = 0; i < dim; i++) s += state[2 * i] * state[2 * i] + state[2 * i + 1] * state[2 * i + 1];
    return s;
}

double orc_prob1(const double *state, int n, int qubit)
{
    const uint64_t dim = 1ull << n;
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (uint64_t i = 0; i < dim; i++)
        if ((i >> qubit) & 1) s += state[2 * i] * state[2 * i] + state[2 * i + 1] * state[2 * i + 1];
    return s;
}

void orc_collapse(double *state, int n, int qubit, int outcome, double p_outcome)
{
    const uint64_t dim = 1ull << n;
    const double scale = 1.0 / sqrt(p_outcome);
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < dim; i++) {
        if ((int)((i >> qubit) & 1) == outcome) { state[2 * i] *= scale; state[2 * i + 1] *= scale; }
        else { state[2 * i] = 0.0; state[2 * i + 1] = 0.0; }
    }
}

double orc_uniform(uint64_t seed, uint64_t counter)
{
    uint64_t z = seed + (counter + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

/* follows apply_measure (:187): probabilities of the qubit, one random draw, projector + renorm */
int orc_measure(double *state, int n, int qubit, uint64_t seed, uint64_t counter)
{
    const double total = orc_norm(state, n);
    const double p1 = orc_prob1(state, n, qubit);
    const double p0 = total - p1;
    const double r = orc_uniform(seed, counter) * total;
    const int outcome = (r >= p0) ? 1 : 0;
    orc_collapse(state, n, qubit, outcome, (outcome ? p1 : p0) / total);
    return outcome;
}

/* follows the static path's final sample_measure (:829, upstream): inverse CDF over |a|^2 */
void orc_sample(const double *state, int n, uint64_t shots, uint64_t seed, uint64_t *out)
{
    const uint64_t dim = 1ull << n;
    /* two-level CDF: block sums of 2^b amplitudes, then a linear walk inside the block */
    const int b = n > 12 ? 12 : n;
    const uint64_t nblk = dim >> b, blk = 1ull << b;
    double *bs = (double *)malloc(sizeof(double) * (nblk + 1));
#pragma omp parallel for schedule(static)
    for (uint64_t k = 0; k < nblk; k++) {
        double s = 0.0;
        for (uint64_t i = k * blk; i < (k + 1) * blk; i++)
            s += state[2 * i] * state[2 * i] + state[2 * i + 1] * state[2 * i + 1];
        bs[k + 1] = s;
    }
    bs[0] = 0.0;
    for (uint64_t k = 0; k < nblk; k++) bs[k + 1] += bs[k];
    const double total = bs[nblk];
#pragma omp parallel for schedule(static)
    for (uint64_t s = 0; s < shots; s++) {
        const double u = orc_uniform(seed, s) * total;
        uint64_t lo = 0, hi = nblk; /* first block with bs[k+1] > u */
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (bs[mid + 1] > u) hi = mid; else lo = mid + 1;
        }
        if (lo >= nblk) lo = nblk - 1;
        double acc = bs[lo];
        uint64_t idx = (lo + 1) * blk - 1;
        for (uint64_t i = lo * blk; i < (lo + 1) * blk; i++) {
            acc += state[2 * i] * state[2 * i] + state[2 * i + 1] * state[2 * i + 1];
            if (acc > u) { idx = i; break; }
        }
        out[s] = idx;
    }
    free(bs);
}
