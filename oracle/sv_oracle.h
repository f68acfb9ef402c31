/*
 * sv_oracle.h -- CPU oracle for the CUNQA state-vector hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the arithmetic that the reference delegates to its
 * third-party CPU simulators (Qiskit-Aer fork 0.17.2 @ 8679dbe6, cunqasimulator v0.1.2; both
 * un-vendored, see /root/reference/CMakeLists.txt:303-311,531-536) at the call sites in
 *   src/backends/simulators/AER/aer_adapters/aer_simulator_adapter.cpp:185-533   (apply_* / measure / reset)
 *   src/backends/simulators/AER/aer_adapters/aer_simulator_adapter.cpp:806-842   (static run + sample)
 * Convention: complex128 array of 2^n amplitudes, qubit 0 = least-significant index bit
 * (Qiskit/Aer little-endian), dense matrices row-major with matrix-index bit m <-> qubits[m].
 *
 * PARITY UNPINNED: the reference's own tests hold no golden vectors / KATs for this path
 * (SURVEY.md 8c), so this oracle is pinned against analytic known answers (GHZ, QFT|x>, U^dagger U)
 * and its numpy twin (oracle/oracle_np.py), not against reference fixtures.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product (cunqa_b200/) never links or calls it.
 */
#ifndef SV_ORACLE_H
#define SV_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* state = interleaved (re,im) doubles, length 2*2^n */
void orc_init_zero(double *state, int n);
/* dense k-qubit gate, optional controls (all must be 1).  mat = row-major 2^k x 2^k (re,im). */
void orc_apply_matrix(double *state, int n, const int *qubits, int k,
                      const int *ctrls, int nc, const double *mat);
/* diagonal k-qubit gate: diag = 2^k (re,im) entries, index bit m <-> qubits[m]. */
void orc_apply_diagonal(double *state, int n, const int *qubits, int k, const double *diag);
/* swap two qubits, optional controls. */
void orc_apply_swap(double *state, int n, int q0, int q1, const int *ctrls, int nc);
/* sum |a|^2 ; and P(qubit==1) */
double orc_norm(const double *state, int n);
double orc_prob1(const double *state, int n, int qubit);
/* project qubit onto outcome and renormalise (Aer apply_measure collapse step). */
void orc_collapse(double *state, int n, int qubit, int outcome, double p_outcome);
/* uniform double in [0,1) from a counter-based generator (splitmix64 of seed/counter);
 * the product uses the same generator so that seeded runs are comparable shot-by-shot. */
double orc_uniform(uint64_t seed, uint64_t counter);
/* measure: r = orc_uniform(seed,counter); outcome = (r >= P0) ; collapses; returns outcome. */
int orc_measure(double *state, int n, int qubit, uint64_t seed, uint64_t counter);
/* sample `shots` basis states from |a|^2 by inverse-CDF: u_s = orc_uniform(seed, s)*norm. */
void orc_sample(const double *state, int n, uint64_t shots, uint64_t seed, uint64_t *out);
/* number of OpenMP threads actually used */
int orc_num_threads(void);
/* set the team size (n <= 0: all online cores), overriding OMP_NUM_THREADS; returns the size in force */
int orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
