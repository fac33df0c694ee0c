/* phare_b200.h — C ABI of libphare_b200.so: the B200 (sm_100a) implementation of PHARE's per-patch
 * hybrid-PIC hot path (SolverPPC's particle sweeps and field solvers).
 *
 * Every entry point cites the reference interface it replaces (paths relative to the PHARE tree).
 * Conventions
 *   - plain C types only; every array pointer is a DEVICE pointer unless its name starts with h_
 *   - every function returns 0 on success, a phb_status otherwise; text via phb_last_error(ctx)
 *   - kernels are enqueued on the context's stream and return without synchronising, except the
 *     functions documented as "host-returning" (they copy a few counters back)
 *   - arrays are C-ordered (last index fastest) with the reference's allocation shapes
 *     (GridLayout::allocSize, src/core/data/grid/gridlayout.hpp:849-864) -> phb_field_shape()
 *   - there is NO CPU fallback: without a CUDA device phb_create fails.
 */
#ifndef PHARE_B200_H
#define PHARE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PHB_OK                = 0,
    PHB_ERR_INVALID       = 1, /* bad argument (dim, interp, null pointer, capacity)            */
    PHB_ERR_CUDA          = 2, /* CUDA runtime error                                            */
    PHB_ERR_MOVE_TWO_CELL = 3, /* a particle moved more than 2 cells in a half push
                                  (boris.hpp:164-165 MoveTwoCellException)                       */
    PHB_ERR_OUTSIDE_GHOST = 4, /* a particle left the ghost box (ion_updater.hpp:256-271, debug)   */
    PHB_ERR_CAPACITY      = 5, /* destination particle array too small                           */
    PHB_ERR_NO_DEVICE     = 6,
    PHB_ERR_PEER_TIMEOUT  = 7  /* phb_peer_wait: a neighbour GPU never signalled                */
} phb_status;

/* HybridQuantity::Scalar subset used on the path (models/quantities/hybrid_quantities.hpp:16-40);
 * centerings follow gridlayout_hybrid_yee.hpp:54-85 */
typedef enum {
    PHB_BX = 0, PHB_BY, PHB_BZ, PHB_EX, PHB_EY, PHB_EZ, PHB_JX, PHB_JY, PHB_JZ,
    PHB_RHO, PHB_VX, PHB_VY, PHB_VZ, PHB_P, PHB_NQTY
} phb_qty;

/* == GridLayout constructor arguments (gridlayout.hpp:122-139) */
typedef struct {
    int      dim;          /* 1,2,3 */
    int      interp;       /* 1,2,3 -> field ghosts {2,4,4}, particle ghosts {1,2,2}
                              (models/options/hybrid_options.hpp:24-25) */
    int      level;        /* AMR level number (used by Ohm spatial hyper-resistivity) */
    int      amr_lower[3]; /* AMR index of the first cell of the patch */
    uint32_t ncells[3];    /* physical cells per direction */
    double   dx[3];
    double   origin[3];
} phb_layout;

typedef struct { double* comp[3]; } phb_vecfield; /* x,y,z components, each allocSize(qty) */

typedef struct { int lower[3], upper[3]; } phb_box; /* inclusive AMR cell box (utilities/box/box.hpp:28) */

/* device-resident SoA particle store; replaces ParticleArray<dim> (particle_array.hpp:21-238)
 * Particle members: particle.hpp:55-61 */
typedef struct {
    int*    icell[3];
    double* delta[3];
    double* v[3];
    double* weight;
    double* charge;
    size_t  n;        /* live particles */
    size_t  capacity; /* allocated slots per column */
} phb_particles;

typedef struct phb_ctx phb_ctx;

/* ---- context / memory ------------------------------------------------------------------- */
int         phb_create(int device, int dim, int interp, phb_ctx** out);
void        phb_destroy(phb_ctx*);
const char* phb_last_error(phb_ctx*);
const char* phb_version(void);
int         phb_set_stream(phb_ctx*, void* cuda_stream); /* cudaStream_t; default: own stream */
void*       phb_get_stream(phb_ctx*);
int         phb_sync(phb_ctx*);
/* 1: products unfused exactly as the reference build (no -march -> no FMA); 0: allow FMA
 * contraction in the gather/Boris arithmetic (positions stay unfused). Default 1. */
int         phb_set_exact(phb_ctx*, int exact);
/* host-returning: synchronises, returns PHB_ERR_MOVE_TWO_CELL / PHB_ERR_OUTSIDE_GHOST raised by a
 * kernel since the last poll, with the offending delta/velocity in phb_last_error; mirrors
 * mpi::any_errors() after moveIons (solver_ppc.hpp:562) */
int         phb_poll_error(phb_ctx*);
/* number of kernels launched by this context so far (bench bookkeeping) */
uint64_t    phb_launch_count(phb_ctx*);

int phb_malloc(phb_ctx*, size_t bytes, void** d_out);
int phb_free(phb_ctx*, void* d);
int phb_memset(phb_ctx*, void* d, int byte, size_t bytes);
int phb_h2d(phb_ctx*, void* d_dst, const void* h_src, size_t bytes);
int phb_d2h(phb_ctx*, void* h_dst, const void* d_src, size_t bytes);
int phb_d2d(phb_ctx*, void* d_dst, const void* d_src, size_t bytes);
int phb_host_alloc(size_t bytes, void** h_out); /* pinned */
int phb_host_free(void* h);

/* ---- geometry (pure host arithmetic) ----------------------------------------------------- */
/* GridLayout::allocSize(qty), gridlayout.hpp:849-864 ; returns element count, fills shape[dim] */
size_t phb_field_shape(const phb_layout*, int qty, uint32_t shape[3]);
int    phb_field_ghosts(int interp);    /* {2,4,4} */
int    phb_particle_ghosts(int interp); /* {1,2,2} */

/* ---- particle store ---------------------------------------------------------------------- */
int phb_particles_alloc(phb_ctx*, size_t capacity, phb_particles* out);
int phb_particles_free(phb_ctx*, phb_particles*);
/* host AoS interop with reference Particle<dim> records (sizeof 56/64/80, particle.hpp:38-73) */
size_t phb_aos_stride(int dim);
int phb_particles_from_aos(phb_ctx*, const void* h_aos, size_t n, phb_particles* dst);
int phb_particles_to_aos(phb_ctx*, const phb_particles* src, void* h_aos);
/* host SoA interop (ContiguousParticles layout, particle_array.hpp:250-354: iCell[n*dim],
 * delta[n*dim], weight[n], charge[n], v[n*3]) */
int phb_particles_from_soa(phb_ctx*, const int* h_icell, const double* h_delta, const double* h_weight,
                           const double* h_charge, const double* h_v, size_t n, phb_particles* dst);
int phb_particles_to_soa(phb_ctx*, const phb_particles* src, int* h_icell, double* h_delta,
                         double* h_weight, double* h_charge, double* h_v);
/* dst[dst_first .. dst_first+count) = src[src_first ..) column by column (device to device);
 * ParticleArray copy / append (ion_updater.hpp:181, std::copy+back_inserter :253) */
int phb_particles_copy(phb_ctx*, const phb_particles* src, size_t src_first, size_t count,
                       phb_particles* dst, size_t dst_first);

/* ---- initial particle loading -------------------------------------------------------
 * MaxwellianParticleInitializer::loadParticles (data/ions/particle_initializers/
 * maxwellian_particle_initializer.hpp:138-199).  The user profile functions are evaluated by the caller at
 * the cell centres (GridLayout::cellCenteredCoordinates) into per-cell arrays, row-major over the patch box
 * (the order of layout.indices(AMRBox)): n[ncell], V[3][ncell], Vth[3][ncell] (, B[3][ncell] for Basis::Magnetic).
 *
 * phb_maxwellian_load_host: parity mode, HOST arrays in and out (ContiguousParticles layout), the
 * reference's sequential std::mt19937_64 algorithm -> the reference's particles bit for bit for a given seed
 * (has_seed = 0: std::random_device, like the reference).  Cells with n < density_cut_off load nothing. */
int phb_maxwellian_load_host(const phb_layout*, const double* h_n, const double* const* h_V,
                             const double* const* h_Vth, const double* const* h_B /* NULL: Basis::Cartesian */,
                             double charge, uint32_t ppc, int has_seed, size_t seed, double density_cut_off,
                             int* h_icell, double* h_delta, double* h_weight, double* h_charge, double* h_v,
                             size_t capacity, size_t* h_count);
/* phb_maxwellian_load: the same distribution sampled on the device with a counter-based generator
 * (Philox4x32-10 keyed by `seed`, counter = (row-major index of the cell in the level's domain box of
 * h_domain_cells[dim] cells) * ppc + i: independent of the patch decomposition; NULL = the patch itself), appended to `out` in cell order.  d_first[ncell+1] = index (relative to the first appended
 * particle) of the first particle of each cell, each cell holding 0 (below the cut-off) or ppc particles;
 * h_total = d_first[ncell].  Cartesian basis.  Statistically equivalent to the reference, not bit-identical. */
int phb_maxwellian_load(phb_ctx*, const phb_layout*, const double* d_n, const phb_vecfield* d_V,
                        const phb_vecfield* d_Vth, const uint32_t* d_first, size_t h_total, double charge,
                        uint32_t ppc, uint64_t seed, const uint32_t* h_domain_cells, phb_particles* out);

/* flat message form of store ranges (ParticlesData::packStream / unpackStream, particles_data.hpp:702-784): a message
 * of `total` particles is column-major, delta[d][total] | v[3][total] | weight[total] | charge[total] | icell[d][total]
 * (phb_particles_flat_bytes bytes); pack copies src[first, first+count) into entries [off, off+count) of every
 * column, unpack appends entries [off, off+count) to dst.  One launch each. */
size_t phb_particles_flat_bytes(int dim, size_t total);
int phb_particles_pack(phb_ctx*, const phb_particles* src, size_t first, size_t count, void* d_buf, size_t total,
                       size_t off);
int phb_particles_unpack(phb_ctx*, const void* d_buf, size_t total, size_t off, size_t count, phb_particles* dst);

/* ---- K1 fused interpolate + Boris push ---------------------------------------------------
 * BorisPusher::move (pusher/boris.hpp:93-138) = prePushStep_ (:180-216), firstSelector,
 * Interpolator::operator()(particle, em, layout) (interpolator.hpp:420-456), accelerate_
 * (:240-300), postPushStep_ (:218-234).  `in` and `out` may be the same store (in-place).
 * first_selector: NULL = noop; else particles whose pre-pushed cell is outside that box keep the
 * pre-pushed position and the old velocity (they are "not selected", ion_updater.hpp:137-140).
 * The second selector is phb_bin / the box predicates of phb_deposit.
 * Pusher::setMeshAndTimeStep (boris.hpp:143-148) is folded in through layout->dx and dt. */
int phb_push(phb_ctx*, const phb_layout*, const phb_vecfield* E, const phb_vecfield* B,
             const phb_particles* in, phb_particles* out, double mass, double dt,
             const phb_box* first_selector);

/* ---- K2 cell binning / counting sort -------------------------------------------------------
 * Replaces CellMap maintenance + partition selectors + erase (cellmap.hpp:411-470,
 * ion_updater.hpp:119-163,245-273).  Key space over `domain` (the patch cell box), G = domain
 * grown by the particle ghost width, and the union `keep[nkeep]` (nonLevelGhostBox,
 * amr_utils.hpp:233-253):
 *    cell in domain                     -> key = rowmajor(cell - domain.lower)        in [0,Nd)
 *    cell in G, in some keep box        -> key = Nd + rowmajor(cell - G.lower)        in [Nd,Nd+Ng)
 *    anything else (level-ghost leavers, or outside G) -> key = Nd+Ng (dropped)
 * `out` receives the particles sorted by key; d_cell_start[k] (Nd+Ng+2 entries) is the index of
 * the first particle of key k.  host-returning: h_counts = {#domain, #patch-ghost, #dropped}.
 * out->n is set to #domain + #patch-ghost. in and out must be distinct stores. */
size_t phb_bin_nkeys(const phb_layout*, const phb_box* domain);
int phb_bin(phb_ctx*, const phb_layout*, const phb_particles* in, phb_particles* out,
            const phb_box* domain, const phb_box* keep, int nkeep, uint32_t* d_cell_start,
            size_t h_counts[3]);
/* phb_bin in three parts, so that the scatter pass can carry the moment deposit (K3+K2 fused):
 *   phb_bin_plan        keys, histogram, slot of every particle of `in`, scan -> d_cell_start (asynchronous);
 *                       the slots stay in the context scratch until phb_deposit_scatter consumes them: no other
 *                       scratch-using call (phb_bin, phb_export*, phb_deposit/phb_push_deposit with a cell_start)
 *                       may come in between
 *   phb_deposit_scatter phb_deposit(in[0,in->n), coef, sel) — cell-ordered kernel over in[0,n_sorted) grouped by
 *                       d_cell_start_old for `domain`, per-particle kernel for the rest — and, in the same pass,
 *                       out[d_cell_start_new[key] + slot] = particle: `out` ends up exactly as phb_bin leaves it
 *                       (updateAndDepositAll_, ion_updater.hpp:245-293: partition/erase + both deposits)
 *   phb_bin_counts      host-returning: h_counts as phb_bin, out->n = #domain + #patch-ghost */
int phb_bin_plan(phb_ctx*, const phb_layout*, const phb_particles* in, const phb_box* domain, const phb_box* keep,
                 int nkeep, uint32_t* d_cell_start);
int phb_deposit_scatter(phb_ctx*, const phb_layout*, const phb_particles* in, size_t n_sorted, double* rho_n,
                        double* rho_q, const phb_vecfield* flux, double coef, const phb_box* sel, int nsel,
                        const phb_box* domain, const uint32_t* d_cell_start_old, const phb_box* keep, int nkeep,
                        phb_particles* out, const uint32_t* d_cell_start_new);
int phb_bin_counts(phb_ctx*, const phb_layout*, const phb_box* domain, const uint32_t* d_cell_start,
                   size_t h_counts[3], phb_particles* out);
/* append to `dst` every particle of src[first,last) whose cell lies in `box` (minus `minus` when
 * not NULL), adding `shift` to its cell: ParticleArray::export_particles (particle_array.hpp
 * :135-160) and ParticlesData pack/unpack with periodic shift (particles_data.hpp:702-784).
 * host-returning: *h_appended. Order of the appended particles = source order. */
int phb_export(phb_ctx*, const phb_layout*, const phb_particles* src, size_t first, size_t last,
               const phb_box* box, const phb_box* minus, const int shift[3], phb_particles* dst,
               size_t* h_appended);

/* the same for `nbox` disjoint boxes in ONE pass (two kernels, one synchronisation): particle k of
 * src[first,last) whose cell lies in boxes[b] is appended to *dsts[b] with iCell += shifts[3b..3b+2].
 * dsts may repeat. This is the whole sender side of fillIonGhostParticles
 * (hybrid_hybrid_messenger_strategy.hpp:410-419) for one patch. host-returning: h_appended[nbox].
 * The order of the appended particles is unspecified. nbox <= 28. */
int phb_export_multi(phb_ctx*, const phb_layout*, const phb_particles* src, size_t first, size_t last, int nbox,
                     const phb_box* boxes, const int* shifts, phb_particles* const* dsts, size_t* h_appended);

/* ---- K3 moment deposit ----------------------------------------------------------------------
 * Interpolator::operator()(range, particleDensity, chargeDensity, flux, layout, coef)
 * (interpolator.hpp:468-504) restricted to particles [first,last) whose cell is in one of
 * `sel[nsel]` (nsel = 0: all particles; this is the `allowed` range the second selector of
 * move() returns, ion_updater.hpp:186-189).
 * d_cell_start != NULL selects the cell-ordered kernel: particles are expected to be (mostly)
 * grouped by the keys of phb_bin for `domain`; particles whose current cell differs from their
 * group's cell are still deposited correctly (through the atomic path). */
int phb_deposit(phb_ctx*, const phb_layout*, const phb_particles*, size_t first, size_t last,
                double* rho_n, double* rho_q, const phb_vecfield* flux, double coef,
                const phb_box* sel, int nsel, const phb_box* domain, const uint32_t* d_cell_start);

/* ---- K1+K3 fused: push and deposit in one pass ------------------------------------------------
 * The body of IonUpdater::updateAndDepositDomain_ / updateAndDepositAll_ for one particle array
 * (ion_updater.hpp:171-219, 228-295): pusher_->move(range, ...) followed by
 * interpolator_(range, density, flux, layout), as ONE pass over parts[first,last): every particle is
 * moved exactly like phb_push (same arithmetic, same first_selector meaning) and, if its NEW cell lies in
 * one of sel[nsel] (nsel = 0: always), deposited exactly like phb_deposit with `coef`.
 *   write_back = 1 : the moved state replaces the stored one (UpdaterMode::all, in place)
 *   write_back = 0 : the store is left untouched (UpdaterMode::domain_only: the reference pushes a
 *                    temporary copy, deposits it and drops it — here the copy is never materialised)
 * d_cell_start/domain as in phb_deposit (cell-ordered kernel; [first,last) must then lie inside the
 * particles of the domain keys, i.e. last <= h_counts[0] of phb_bin); NULL = any order.
 * Errors (move of more than two cells) are raised through phb_poll_error like phb_push. */
int phb_push_deposit(phb_ctx*, const phb_layout*, const phb_vecfield* E, const phb_vecfield* B,
                     phb_particles* parts, size_t first, size_t last, double mass, double dt,
                     const phb_box* first_selector, int write_back, double* rho_n, double* rho_q,
                     const phb_vecfield* flux, double coef, const phb_box* sel, int nsel,
                     const phb_box* domain, const uint32_t* d_cell_start);

/* Interpolator::operator()(Particle&, Electromag const&, GridLayout const&) (interpolator/interpolator.hpp:420-456) on its
 * own: E and B interpolated at the position of parts[first,last), d_eb[6 * (i - first) + {0..5}] = {Ex,Ey,Ez,Bx,By,Bz}.
 * The same device functions phb_push inlines (same bits). */
int phb_gather(phb_ctx*, const phb_layout*, const phb_vecfield* E, const phb_vecfield* B, const phb_particles* parts,
               size_t first, size_t last, double* d_eb);

/* phb_push_plan: the first half of IonUpdater::updateAndDepositAll_ for the domain array (ion_updater.hpp:228-245) in ONE
 * pass over the store: pusher_->move in place (== phb_push(parts, parts, first = NULL), same bits) and, while the moved
 * particle is still in registers, the count half of the re-binning (== phb_bin_plan(parts, domain, keep)): key of the new
 * cell, slot inside it, histogram -> d_cell_start.  phb_deposit_scatter consumes the plan exactly as after phb_bin_plan.
 * d_cell_start_old / n_sorted (may be NULL / 0): the current ordering of parts[0, n_sorted) by the keys of phb_bin for
 * `domain`; with it the ordered part goes through the strip kernel (csrc/strip.cuh: the E,B nodes of each strip of cells
 * and their stencil halo staged in shared memory by bulk copies, gathers from shared memory), the rest and every store
 * without an ordering through the streaming kernel (E,B through L1). */
int phb_push_plan(phb_ctx*, const phb_layout*, const phb_vecfield* E, const phb_vecfield* B, phb_particles* parts,
                  size_t n_sorted, double mass, double dt, const phb_box* domain, const uint32_t* d_cell_start_old,
                  const phb_box* keep, int nkeep, uint32_t* d_cell_start);

/* ---- cell-ordered passes with the E,B nodes of a block of cells staged in shared memory (csrc/tile.cuh) -----------
 * The kernels behind these entry points (and behind the cell-ordered path of phb_push_deposit) give one CTA a compact
 * block of cells, stage the block's E,B nodes + stencil halo ONCE in shared memory with bulk asynchronous copies (TMA)
 * and serve every gather from there.  `domain` / d_cell_start as in phb_deposit: parts[0, n_sorted) is ordered by the
 * keys of phb_bin for `domain`; parts[n_sorted, n) (received since the last binning) may be in any order.
 *
 * phb_push_cells: BorisPusher::move (pusher/boris.hpp:93-138) exactly like phb_push with first_selector = NULL
 * (same arithmetic, same bits) through the strip kernel (csrc/strip.cuh: E,B of each strip of cells staged in shared
 * memory by bulk copies).  `out` may be `in`, or a store whose weight/charge columns alias in's. */
int phb_push_cells(phb_ctx*, const phb_layout*, const phb_vecfield* E, const phb_vecfield* B, const phb_particles* in,
                   phb_particles* out, size_t n_sorted, double mass, double dt, const phb_box* domain,
                   const uint32_t* d_cell_start);
/* phb_push_deposit_plan: IonUpdater::updateAndDepositAll_ for the domain array (ion_updater.hpp:228-295) in one pass
 * over the store: pusher_->move in place, the deposit of the moved particles whose new cell lies in sel[nsel]
 * (= phb_push_deposit with write_back = 1), and the bookkeeping of the partition / erase that follows (:245-273):
 * per cell the number of particles that stay and, for the few that change cell, their rank among the arrivals of the
 * new cell; then the scan -> d_cell_start_new (asynchronous; the same values phb_bin / phb_bin_plan produce).
 * phb_scatter_planned then moves the data: `out` ends up exactly as phb_bin(parts, out, ...) leaves it (same
 * d_cell_start, same per-cell multisets); phb_bin_counts returns the class counts.  The plan lives in the context
 * until phb_scatter_planned consumes it. */
int phb_push_deposit_plan(phb_ctx*, const phb_layout*, const phb_vecfield* E, const phb_vecfield* B,
                          phb_particles* parts, size_t n_sorted, double mass, double dt, double* rho_n, double* rho_q,
                          const phb_vecfield* flux, double coef, const phb_box* sel, int nsel, const phb_box* domain,
                          const uint32_t* d_cell_start_old, const phb_box* keep, int nkeep, uint32_t* d_cell_start_new);
int phb_scatter_planned(phb_ctx*, const phb_layout*, const phb_particles* in, size_t n_sorted, const phb_box* domain,
                        const uint32_t* d_cell_start_old, const phb_box* keep, int nkeep, phb_particles* out,
                        const uint32_t* d_cell_start_new);

/* ---- predicted re-binning (csrc/predict.cu): the `all` sweep of IonUpdater::updateAndDepositAll_
 * (ion_updater.hpp:228-295) in ONE pass over the store -----------------------------------------------------------
 * A PPC step pushes every particle twice from the same state (solver_ppc.hpp:325-333); the cell it ends in after the
 * second push is the one the first push predicts, except for the few particles whose predicted position lies within
 * eps (2^-12 of a cell; PHB_PREDICT_EPS) of a cell face.
 *   phb_push_deposit_predict  == phb_push_deposit(parts[0,n), write_back = 0) for a store whose first n_sorted particles
 *                             are ordered by d_cell_start (phb_bin keys of `domain`), and the plan of the re-binning into
 *                             [domain | patch ghosts kept by keep[] | erased]: rank of every stayer in its cell, rank of
 *                             every mover among the arrivals of its new cell, list of the face-near particles.  The plan
 *                             lives in the caller's device buffer d_plan (phb_predict_plan_bytes(L, domain,
 *                             parts->capacity) bytes) until phb_push_deposit_rebin consumes it; `parts` must not change
 *                             in between.
 *   phb_push_deposit_rebin    pusher_->move + deposit + partition/erase: the face-near particles are moved with THESE
 *                             fields and join the plan, scan -> d_cell_start_new, then every particle of `in` is moved,
 *                             deposited (sel[]) and written to the slot its plan reserved in `out`.  Every plan is checked
 *                             against the particle's actual new cell.
 *   phb_predict_counts        host-returning: h_counts[0..2] as phb_bin_counts, h_counts[3] = plans that did not hold
 *                             ("misfiled": such a particle is pushed and deposited correctly but filed under the predicted
 *                             cell).  0 -> `out` is exactly what phb_bin leaves; > 0 -> the caller restores the order with
 *                             phb_bin(out -> ...) before using d_cell_start_new.
 * Only for (dim, interp) whose cell support fits the tile kernel (phb_predict_supported). */
/* eps of the predicted re-binning (default 2^-12 of a cell; the environment variable PHB_PREDICT_EPS overrides both).  A host that
 * sees failed plans (phb_predict_counts) raises it: the two sweeps of its steps differ by more than the default assumes. */
int    phb_set_predict_eps(phb_ctx*, double eps);
double phb_get_predict_eps(phb_ctx*);
int    phb_predict_supported(const phb_layout*);
size_t phb_predict_plan_bytes(const phb_layout*, const phb_box* domain, size_t capacity);
int phb_push_deposit_predict(phb_ctx*, const phb_layout*, const phb_vecfield* E, const phb_vecfield* B,
                             const phb_particles* parts, size_t n_sorted, double mass, double dt, double* rho_n,
                             double* rho_q, const phb_vecfield* flux, double coef, const phb_box* sel, int nsel,
                             const phb_box* domain, const uint32_t* d_cell_start, const phb_box* keep, int nkeep,
                             void* d_plan, size_t plan_bytes);
int phb_push_deposit_rebin(phb_ctx*, const phb_layout*, const phb_vecfield* E, const phb_vecfield* B,
                           const phb_particles* in, size_t n_sorted, double mass, double dt, double* rho_n, double* rho_q,
                           const phb_vecfield* flux, double coef, const phb_box* sel, int nsel, const phb_box* domain,
                           const uint32_t* d_cell_start_old, const phb_box* keep, int nkeep, phb_particles* out,
                           uint32_t* d_cell_start_new, void* d_plan, size_t plan_bytes);
int phb_predict_counts(phb_ctx*, const phb_layout*, const phb_box* domain, const uint32_t* d_cell_start,
                       const void* d_plan, size_t h_counts[4], phb_particles* out);

/* ---- GridLayout primitives on their own (test support for SURVEY 8 row a14) ---------------------------------------
 * The device functions Faraday / Ampere / Ohm are built from, applied to one array of quantity `qty` (allocation shape
 * of phb_field_shape):
 *   op 0  GridLayout::deriv(field, index, direction = arg)   gridlayout.hpp:571-630.  d_out has the allocation of the
 *         other centering along `arg` (allocSizeDerived, :866-880) and is written on that centering's physical range
 *   op 1  GridLayout::laplacian(field, index)                gridlayout.hpp:640-700, written on the physical box of qty
 *   op 2  the Yee linear combinations (momentsToEx, BzToEx, JxToMoments, ... gridlayout_hybrid_yee.hpp:350-749) by kind
 *         per direction, arg = kx | ky << 2 | kz << 4 with 0 = same centering, 1 = primal -> dual {0,+1}, 2 = dual -> primal
 *         {-1,0}; d_out has the shape of d_in and is written wherever the stencil stays inside the array. */
int phb_gridlayout_probe(phb_ctx*, const phb_layout*, int op, int qty, int arg, const double* d_in, double* d_out);

/* ---- particle splitting (level refinement) ---------------------------------------
 * ParticlesRefineOperator::refine_ (amr/data/particles/refine/particles_data_split.hpp:142-231) for one source
 * array: every coarse particle of coarse[first,last) is moved to the fine index space (toFineGrid, :32-46), and if
 * it lies within max_cell_distance cells of a destination box it is split with the (delta, weight) table of
 * Splitter<dim, interp, nbRefinedPart> (splitter.hpp:71-106; h_deltas[nref*dim], h_weights[nref], float32 like the
 * reference); the refined particles whose cell lies in that box are appended to `fine`.  Boxes are fine-level AMR
 * cell boxes and must be disjoint.  host-returning: *h_appended.  Output order: source order, then pattern order. */
int phb_split(phb_ctx*, const phb_particles* coarse, size_t first, size_t last, int nref, const float* h_deltas,
              const float* h_weights, int max_cell_distance, const phb_box* fine_boxes, int nbox,
              phb_particles* fine, size_t* h_appended);

/* ---- K4-K7 field solvers and pointwise ops -------------------------------------------------- */
/* Faraday::operator()(B,E,Bnew,dt)  faraday/faraday.hpp:28-97 */
int phb_faraday(phb_ctx*, const phb_layout*, const phb_vecfield* B, const phb_vecfield* E,
                phb_vecfield* Bnew, double dt);
/* Ampere::operator()(B,J)           ampere/ampere.hpp:26-95 */
int phb_ampere(phb_ctx*, const phb_layout*, const phb_vecfield* B, phb_vecfield* J);
/* Ohm::operator()(n,Ve,Pe,B,J,Enew) ohm/ohm.hpp:48-270 ; hyper_mode 0 = constant, 1 = spatial */
int phb_ohm(phb_ctx*, const phb_layout*, const double* n, const phb_vecfield* Ve, const double* Pe,
            const phb_vecfield* B, const phb_vecfield* J, phb_vecfield* Enew, double eta, double nu,
            int hyper_mode);
/* Electrons::update: Ve = Vi - project(J)/Ne on the physical primal box; Pe = Ne*Te everywhere
 * (data/electrons/electrons.hpp:102-128,202-212,304-314) */
int phb_electrons_update(phb_ctx*, const phb_layout*, const double* Ne, const phb_vecfield* Vi,
                         const phb_vecfield* J, double Te, phb_vecfield* Ve, double* Pe);
/* Ions::computeChargeDensity / computeMassDensity / computeBulkVelocity (data/ions/ions.hpp:75-145)
 * h_* arrays of npop device pointers / masses live on the host */
/* rho_q_tot, rho_m_tot and V may each be NULL (Ions::computeChargeDensity / computeMassDensity / computeBulkVelocity called
 * separately, ions.hpp:75,92,111); at least one must be given */
int phb_ions_totals(phb_ctx*, size_t nnodes, int npop, const double* const* h_rho_n,
                    const double* const* h_rho_q, const phb_vecfield* h_flux, const double* h_mass,
                    double* rho_q_tot, double* rho_m_tot, phb_vecfield* V);
/* core::average (utilities/algorithm.hpp:68-77): avg = (a+b)*0.5 over n doubles */
int phb_average(phb_ctx*, size_t n, const double* a, const double* b, double* avg);
/* count <= 8 averages in one launch: avg[k][i] = (a[k][i] + b[k][i]) / 2 for i < n[k] (the six components of average_) */
int phb_average_many(phb_ctx*, int count, const size_t* n, const double* const* a, const double* const* b,
                     double* const* avg);

/* ---- K8 same-level periodic halo on one device ----------------------------------------------
 * dst region = box `dst_box` (local array indices, inclusive) of array `dst` with shape
 * dst_shape; src likewise; op 0 = copy (ghost fill, field_data.hpp:212-290), 1 = += (border sum,
 * field_operate_transaction.hpp:33 PlusEquals), 2 = max (SetMax, types.hpp:577-581) */
int phb_box_op(phb_ctx*, int dim, double* dst, const uint32_t dst_shape[3], const uint32_t dst_lo[3],
               const double* src, const uint32_t src_shape[3], const uint32_t src_lo[3],
               const uint32_t extent[3], int op);
/* pack / unpack a box to / from a contiguous buffer (NCCL send / recv staging) */
int phb_box_pack(phb_ctx*, int dim, const double* src, const uint32_t src_shape[3],
                 const uint32_t src_lo[3], const uint32_t extent[3], double* buf);
int phb_box_unpack(phb_ctx*, int dim, double* dst, const uint32_t dst_shape[3],
                   const uint32_t dst_lo[3], const uint32_t extent[3], const double* buf, int op);

/* batched form: one launch executes `nops` box operations (a whole exchange phase).  Descriptors are
 * in DEVICE memory; shapes/lows/extents are padded with 1/0/1 in unused trailing directions;
 * `first` = running sum of the element counts of the preceding descriptors. Overlapping
 * destinations inside one batch are only allowed for op 0 with identical values and op 2.
 * op 3 = fill with quiet NaN (setNaNsOnFieldGhosts of a refined level; src is ignored). */
typedef struct {
    double*       dst;
    const double* src;
    uint32_t      dst_shape[3], dst_lo[3], src_shape[3], src_lo[3], ext[3];
    int32_t       op;
    uint64_t      first;
} phb_box_desc;
int phb_box_op_batch(phb_ctx*, const phb_box_desc* d_ops, int nops, uint64_t total_elements);

/* ---- coarse <-> fine level operators (refinement ratio 2) ---------------------------
 * A phb_field_view is an array of one quantity together with the AMR field index of its first element
 * (primal directions count nodes, dual directions count cells), i.e. what the reference passes to its
 * refiners / coarseners as (Field, ghost field box): local index = AMR index - lo (AMRToLocal,
 * amr/resources_manager/amr_utils.hpp).  Boxes are inclusive AMR FIELD-index boxes of `qty`. */
typedef struct { double* data; uint32_t shape[3]; int32_t lo[3]; } phb_field_view;
typedef enum {
    PHB_REFINE_DEFAULT       = 0, /* DefaultFieldRefiner: linear, writes NaN nodes only (field_refiner.hpp:32-170,
                                     field_linear_refine.hpp:29-129, linear_weighter.cpp:9-56)                        */
    PHB_REFINE_MAGNETIC      = 1, /* MagneticFieldRefiner: coarse faces copied, NaN nodes only
                                     (magnetic_field_refiner.hpp:28-195)                                              */
    PHB_REFINE_MAGNETIC_INIT = 2, /* MagneticFieldInitRefiner: same, unconditional (magnetic_field_init_refiner.hpp)  */
    PHB_REFINE_ELECTRIC      = 3  /* ElectricFieldRefiner: edge-conserving, NaN nodes only
                                     (electric_field_refiner.hpp:30-345)                                              */
} phb_refine_op;
typedef enum {
    PHB_COARSEN_ELECTRIC = 0, /* ElectricFieldCoarsener (electric_field_coarsener.hpp:38-150) */
    PHB_COARSEN_MOMENTS  = 1  /* MomentsCoarsener: injection (moments_coarsener.hpp:30-84)     */
} phb_coarsen_op;
/* FieldRefineOperator::refine for one destination box (field_refine_operator.hpp:60-102): every fine index of
 * `fine_box` gets refiner(coarse, fine, index).  The coarse view must hold coarsen(fine_box) (+1 for the linear and
 * electric stencils). */
int phb_field_refine(phb_ctx*, int dim, int op, int qty, const phb_field_view* coarse,
                     const phb_field_view* fine, const phb_box* fine_box);
/* MagneticRefinePatchStrategy::postprocessRefine (magnetic_refine_patch_strategy.hpp:66-125, 1-D :134-141,
 * 2-D :143-190, 3-D :192-372): the NEW fine faces (odd index along the component's own direction) of the field boxes
 * of the cell box `fine_cell_box` get the divergence-free Toth-Roe value from the coarse faces around them.
 * Faces whose cell lies in one of the `excluded` cell boxes (<= 27; the patches of the level) are left alone, so one
 * call covers every fill box of a patch's level-ghost layer. */
int phb_magnetic_postprocess(phb_ctx*, const phb_layout* fine, const phb_vecfield* B, const phb_box* fine_cell_box,
                             const phb_box* excluded, int nexcluded);
/* FieldCoarsenOperator::coarsen for one coarse box (field_coarsen_operator.hpp:95-121) */
int phb_field_coarsen(phb_ctx*, int dim, int op, int qty, const phb_field_view* fine,
                      const phb_field_view* coarse, const phb_box* coarse_box);
/* setNaNsOnFieldGhosts (hybrid_hybrid_messenger_strategy.hpp:924-955) / zero(): dst[box] = value */
int phb_box_fill(phb_ctx*, int dim, double* dst, const uint32_t dst_shape[3], const uint32_t dst_lo[3],
                 const uint32_t extent[3], double value);
/* core::operate<PlusEqualsProduct> (utilities/algorithm.hpp:166-183, types.hpp:584-588): dst += src * coef;
 * SolverPPC::accumulateFluxSum (solver_ppc.hpp:263-276) */
int phb_axpy(phb_ctx*, size_t n, double* dst, const double* src, double coef);

/* ---- peer-memory halo exchange over NVLink (one process per GPU, GPUs of one node) ---------------
 * The receive areas of the field exchange phases live in ONE device allocation per rank (phb_malloc) that
 * every other rank maps through CUDA IPC; phb_box_op_batch descriptors may then name a peer's memory as
 * destination (the pack kernel stores straight into the neighbour GPU).  Ordering between the two GPUs:
 *   phb_peer_signal(flags, v): after all work already enqueued on the stream, *flags[i] = v[i] (release.sys);
 *                              flags[i] normally point into peer memory
 *   phb_peer_wait(flags, v)  : the stream waits until every *flags[i] >= v[i] (acquire.sys); gives up after
 *                              timeout_s -> PHB_ERR_PEER_TIMEOUT through phb_poll_error
 * n <= 32 flag words per call. */
int phb_ipc_export(phb_ctx*, void* d_ptr, unsigned char h_handle[64]);
int phb_ipc_open(phb_ctx*, const unsigned char h_handle[64], void** d_peer_ptr);
int phb_ipc_close(phb_ctx*, void* d_peer_ptr);
int phb_peer_signal(phb_ctx*, int n, uint64_t* const* h_flag_ptrs, const uint64_t* h_values);
int phb_peer_wait(phb_ctx*, int n, uint64_t* const* h_flag_ptrs, const uint64_t* h_values, double timeout_s);

/* One whole exchange phase in one call, with DEVICE-side phase counters: pack (K8 batch, remote stores into the neighbours'
 * receive areas) -> signal -> local box ops -> wait -> unpack.  The value a signal publishes / a wait expects is
 * *counter + 1, advanced by the kernel itself (counters are device words of the caller, one per peer and direction), so
 * the descriptor of a phase is built once and never changes.  Replaces the SAMRAI RefineSchedule::fillData of one
 * communicator (hybrid_hybrid_messenger_strategy.hpp:376-497). */
typedef struct {
    const phb_box_desc* pre;   int n_pre;   uint64_t total_pre;   /* device tables as for phb_box_op_batch (may be empty) */
    const phb_box_desc* local; int n_local; uint64_t total_local;
    const phb_box_desc* post;  int n_post;  uint64_t total_post;
    int       n_signal;  uint64_t* signal_flag[32]; uint64_t* signal_counter[32]; /* flag: device word in the PEER's arena */
    int       n_wait;    uint64_t* wait_flag[32];   uint64_t* wait_counter[32];   /* flag: device word in MY arena        */
    double    timeout_s;
} phb_peer_phase_desc;
int phb_peer_phase(phb_ctx*, const phb_peer_phase_desc* h_phase);


#ifdef __cplusplus
}
#endif
#endif
