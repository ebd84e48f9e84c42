/* oracle/port_pvq.h -- TEST INFRASTRUCTURE ONLY (see port.h). */
#ifndef DAALA_ORACLE_PORT_PVQ_H
#define DAALA_ORACLE_PORT_PVQ_H
#include <stdint.h>
#include "port.h"

int port_pvq_cos(int32_t x);
int port_pvq_sin(int32_t x);
int port_vector_log_mag(const od_coeff *x, int n);
int port_compute_householder(int16_t *r, int n, int32_t gr, int *sign, int shift);
void port_apply_householder(int16_t *out, const int16_t *x, const int16_t *r, int n);
int32_t port_gain_expand(int32_t cg0, int q0, int beta);
int32_t port_pvq_compute_gain(const int16_t *x, int n, int q0, int32_t *g, int beta, int bshift);
int port_pvq_compute_max_theta(int32_t qcg, int beta);
int32_t port_pvq_compute_theta(int t, int max_theta);
int port_pvq_compute_k(int32_t qcg, int itheta, int noref, int n, int beta);
void port_pvq_synthesis_partial(od_coeff *xcoeff, const od_coeff *ypulse, const int16_t *r16, int n,
 int noref, int32_t g, int32_t theta, int m, int s, const int16_t *qm_inv);
double port_pvq_search_rdo_double(const int16_t *xcoeff, int n, int k, od_coeff *ypulse, double g2,
 double pvq_norm_lambda, int prev_k);
double port_pvq_rate(int qg, int icgr, int theta, int ts, const od_coeff *y0, int k, int n,
 int is_keyframe, int pli);
int port_pvq_theta(od_coeff *out, const od_coeff *x0, const od_coeff *r0, int n, int q0, od_coeff *y,
 int *itheta, int *max_theta, int *vk, int beta, double *skip_diff, int is_keyframe, int pli,
 const int16_t *qm, const int16_t *qm_inv, double pvq_norm_lambda);
#endif
