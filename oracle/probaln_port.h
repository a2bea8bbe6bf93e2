/* oracle/probaln_port.h -- TEST INFRASTRUCTURE ONLY.  See probaln_port.c. */
#ifndef ORACLE_PROBALN_PORT_H
#define ORACLE_PROBALN_PORT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Restated htslib-1.17 probaln_glocal with the parameter struct unpacked.
 * state[l_query], q[l_query] as htslib; optional extras for tests:
 *   s_out[l_query+2]  per-row scaling factors,
 *   pmax_out[l_query] normalised max posterior per row (the argument of log(1-max)),
 *   pb_out            b[0][0] (== 1 up to rounding).
 * Returns the Phred-scaled likelihood like htslib, INT_MIN on allocation failure, 0 on empty input. */
int oracle_probaln_glocal_ex(const uint8_t *ref, int l_ref, const uint8_t *query, int l_query,
                             const uint8_t *iqual, float par_d, float par_e, int par_bw,
                             int *state, uint8_t *q, double *s_out, double *pmax_out, double *pb_out);

/* Band cells of one instance: sum_i (min(Lr,i+bw) - max(1,i-bw) + 1), bw as the HMM derives it. */
long oracle_probaln_cells(int l_ref, int l_query, int c_bw);

typedef void (*oracle_hmm_trace_fn)(void *ud, const uint8_t *ref, int l_ref, const uint8_t *query, int l_query,
                                    const uint8_t *iqual, float d, float e, int bw, const int *state,
                                    const uint8_t *q);
/* Per-thread hook invoked after every probaln_glocal() call made through the htslib symbol. */
void oracle_probaln_set_trace(oracle_hmm_trace_fn fn, void *ud);

#ifdef __cplusplus
}
#endif
#endif
