/* TEST INFRASTRUCTURE — the CPU oracle.  NOT part of the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * A plain-C restatement of PHARE's per-patch hybrid-PIC arithmetic (see phare_oracle.c for the
 * reference file:line each function follows).  Parity is PINNED: tests/test_oracle_vs_reference.py
 * checks every function bit-for-bit against the reference's own headers compiled in place
 * (oracle/ref/ref_driver.cpp -> oracle/_ref/libphare_ref.so) and against the reference's golden
 * generators (tests/golden/).  Same structs as the product C ABI (include/phare_b200.h) but every
 * pointer is a HOST pointer.  Must be built with -ffp-contract=off (no FMA), like the reference
 * default build.
 */
#ifndef PHARE_ORACLE_H
#define PHARE_ORACLE_H
#include "../include/phare_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

size_t pho_field_shape(const phb_layout*, int qty, uint32_t shape[3]);

/* returns 0, or PHB_ERR_MOVE_TWO_CELL (err_delta/err_vel receive the offending values) */
int pho_push(const phb_layout*, const phb_vecfield* E, const phb_vecfield* B, const phb_particles* in,
             phb_particles* out, double mass, double dt, const phb_box* first_selector,
             double* err_delta, double* err_vel);
/* interpolated E,B at each particle (for unit tests of the gather): eb[n][6] */
int pho_gather(const phb_layout*, const phb_vecfield* E, const phb_vecfield* B, const phb_particles* in,
               double* eb);
int pho_deposit(const phb_layout*, const phb_particles*, size_t first, size_t last, double* rho_n,
                double* rho_q, const phb_vecfield* flux, double coef, const phb_box* sel, int nsel);
size_t pho_bin_nkeys(const phb_layout*, const phb_box* domain);
int pho_bin(const phb_layout*, const phb_particles* in, phb_particles* out, const phb_box* domain,
            const phb_box* keep, int nkeep, uint32_t* cell_start, size_t counts[3]);
int pho_export(const phb_layout*, const phb_particles* src, size_t first, size_t last, const phb_box* box,
               const phb_box* minus, const int shift[3], phb_particles* dst, size_t* appended);

int pho_faraday(const phb_layout*, const phb_vecfield* B, const phb_vecfield* E, phb_vecfield* Bnew,
                double dt);
int pho_ampere(const phb_layout*, const phb_vecfield* B, phb_vecfield* J);
int pho_ohm(const phb_layout*, const double* n, const phb_vecfield* Ve, const double* Pe,
            const phb_vecfield* B, const phb_vecfield* J, phb_vecfield* Enew, double eta, double nu,
            int hyper_mode);
int pho_electrons_update(const phb_layout*, const double* Ne, const phb_vecfield* Vi, const phb_vecfield* J,
                         double Te, phb_vecfield* Ve, double* Pe);
int pho_ions_totals(size_t nnodes, int npop, const double* const* rho_n, const double* const* rho_q,
                    const phb_vecfield* flux, const double* mass, double* rho_q_tot, double* rho_m_tot,
                    phb_vecfield* V);
int pho_average(size_t n, const double* a, const double* b, double* avg);
int pho_box_op(int dim, double* dst, const uint32_t dst_shape[3], const uint32_t dst_lo[3],
               const double* src, const uint32_t src_shape[3], const uint32_t src_lo[3],
               const uint32_t extent[3], int op);

/* coarse <-> fine level operators (see include/phare_b200.h for the reference lines) */
int pho_field_refine(int dim, int op, int qty, const phb_field_view* coarse, const phb_field_view* fine,
                     const phb_box* fine_box);
int pho_magnetic_postprocess(const phb_layout* fine, const phb_vecfield* B, const phb_box* fine_cell_box,
                             const phb_box* excluded, int nexcluded);
int pho_field_coarsen(int dim, int op, int qty, const phb_field_view* fine, const phb_field_view* coarse,
                      const phb_box* coarse_box);
int pho_box_fill(int dim, double* dst, const uint32_t dst_shape[3], const uint32_t dst_lo[3], const uint32_t extent[3],
                 double value);
int pho_axpy(size_t n, double* dst, const double* src, double coef);

/* B-spline weights of one direction (Weighter<order>::computeWeight + computeStartLeftShift,
 * interpolator.hpp:54-125,519-549): returns the start index; w[order+1] */
int pho_weights(int order, int dual, uint32_t local_cell, double delta, double* w);

#ifdef __cplusplus
}
#endif
#endif
