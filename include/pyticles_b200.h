/* pyticles_b200 -- C ABI of the B200 (sm_100a) SPH step hot path.
 *
 * This is the drop-in boundary for pyticles' hot path.  The reference has no FFI of its
 * own: the boundary there is Python duck typing, `import c_forces as forces`
 * (run_scripts/ospana.py:13) / `import f_properties as properties` (particles.py:28), with
 * Cython (`pairsep.pyx`, `c_forces.pyx`) and f2py modules behind it.  Each entry point
 * below names the reference routine it replaces (file:line into the reference tree).
 * INTEGRATION.md shows the ctypes stub a pyticles maintainer would add.
 *
 * Conventions
 *   - Plain C: pointers, sizes and POD structs only; no C++/torch types.
 *   - Every pointer named d_* (and every pointer inside sph_buffers) is a DEVICE pointer
 *     owned by the caller (the Python side allocates them as torch tensors); the library
 *     never allocates or frees device memory and keeps no state between calls.
 *   - `stream` is a cudaStream_t passed as void*.  All work is enqueued on it; no call
 *     synchronises the device.  Results that the host needs (pair count, overflow,
 *     rebuild decision) are left in the device-side `sph_status`, which the caller copies
 *     back when it wants them.
 *   - Return value: SPH_OK (0), or a negative SPH_E_* for bad arguments, or a positive
 *     cudaError_t from the launch.  Nothing throws.
 *   - Particle arrays keep pyticles' layout: r, v, vdot are C-order [n,3] float64;
 *     m, h, t, rho, p, pco, u, udot are [n] float64 (particles.py:122-127,323-343).
 */
#ifndef PYTICLES_B200_H
#define PYTICLES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPH_ABI_VERSION 4      /* sph_version() reports the same number */

enum {
    SPH_OK = 0,
    SPH_E_BADARG = -1,    /* null pointer, n < 0, capacity <= 0, ... */
    SPH_E_GRID = -2,      /* box / cutoff cannot be gridded (non-positive or non-finite) */
    SPH_E_TOOBIG = -3     /* more cells or particles than 32-bit indices allow */
};

/* status flags (sph_status.flags) */
enum {
    SPH_F_NONFINITE = 1,      /* a position is NaN/Inf */
    SPH_F_OUT_OF_RANGE = 2,   /* a position is outside [-L/4, 5L/4]: the fp32 pre-filter is
                                 switched off and every candidate is tested in fp64 */
    SPH_F_OUT_OF_BOX = 4,     /* a position is outside [0, L): interior cells keep the
                                 minimum-image test */
    SPH_F_NBR_OVERFLOW = 8,   /* some particle has more neighbours than `max_nbrs`;
                                 sph_status.max_count says how many are needed */
    SPH_F_OUT_OF_SLAB = 16,   /* a particle lies outside the local cell-layer range */
    SPH_F_TILE_FALLBACK = 32, /* the cell-group (tile) neighbour kernel met a case outside its
                                 fixed capacities (a cell with > 64 particles, > 1280 particles
                                 in the 64 cells around a group, > 32 hits in one stream's list,
                                 positions far outside the box): the general kernel did the
                                 pass.  The neighbour structure is the same either way (rows
                                 hold the same sets); only speed differs. */
    SPH_F_HALO_OVERFLOW = 64  /* a boundary cell layer holds more particles than the halo buffers
                                 (sph_status.halo_count says how many): the ghosts are incomplete */
};

/* Device-resident status block (64 bytes).  Zero it with sph_status_reset before a build. */
typedef struct sph_status {
    uint32_t flags;
    uint32_t max_count;          /* with SPH_F_NBR_OVERFLOW: the largest per-particle neighbour count (the capacity
                                  * needed); otherwise a lower bound of it, possibly 0 */
    uint32_t halo_count[2];      /* particles in the left / right boundary cell layer (sph_cells_begin) */
    uint32_t ghost_count[2];     /* ghosts received from the left / right neighbour (sph_halo_unpack) */
    unsigned long long dsq_max_bits; /* ponder_rebuild: max |r_old - r|^2 as ordered bits */
    uint32_t rebuild;            /* ponder_rebuild result (neighbour_list.py:225-234) */
    uint32_t reserved[7];
} sph_status;

/* Cell grid description (host POD, passed by pointer, copied into kernel arguments). */
typedef struct sph_grid {
    double box[3];       /* xmax, ymax, zmax   (box.py:19-23) */
    double thr;          /* cutoff^2 + tolerance^2   (neighbour_list.py:155-157,178) */
    double w[3];         /* cell width per dimension (>= sqrt(thr) * (1 + 2^-20)) */
    double inv_w[3];     /* nc / box */
    int32_t nc[3];       /* cells per dimension across the whole periodic box */
    int32_t lo[3];       /* first global cell layer of the local grid (0 on one GPU) */
    int32_t ncl[3];      /* local cell layers (== nc on one GPU) */
    int32_t wrap[3];     /* 1: local grid is periodic in this dimension */
    uint32_t mask[3];    /* bit positions of each dimension in the in-block Morton code */
    uint32_t ncode;      /* number of cell codes = nblk[0]*nblk[1]*nblk[2] << lbits */
    float thr_in;        /* fp32 pre-filter: rsq32 <  thr_in  => certainly inside  */
    float thr_out;       /*                  rsq32 >= thr_out => certainly outside */
    uint32_t top[3];     /* in-block code bits of the last local layer of each dimension */
    uint32_t magic0;     /* ceil(2^32 / nblk[0]): multiply-high division of the block index */
    /* cell code = (block index << lbits) | in-block Morton code; blocks of 2^lb layers per
     * dimension (lb <= 3), numbered row-major (x fastest) */
    uint32_t lb[3];
    uint32_t nblk[3];
    uint32_t lbits;
    uint32_t magic1;     /* ceil(2^32 / nblk[1]) */
} sph_grid;

/* Equation of state constants (properties.py:18-20; feos.eos.{adash,bdash,kbdash}). */
typedef struct sph_eos {
    double adash, bdash, kbdash;
} sph_eos;

/* Caller-owned device buffers of one neighbour structure.  Sizes in elements. */
typedef struct sph_buffers {
    int32_t n;              /* particles */
    int32_t max_nbrs;       /* ELL capacity per particle (multiple of 4) */
    /* cell list */
    uint32_t *cell_count;   /* [ncode + 1]  (the last cell collects the unused ghost slots, see n_valid) */
    uint32_t *cell_start;   /* [ncode + 2]  */
    uint32_t *scan_tmp;     /* [sph_scan_tmp_elems(ncode + 1)] */
    uint32_t *code;         /* [n] cell code of particle i (original order) */
    uint32_t *rank;         /* [n] arrival rank inside its cell */
    int32_t *perm;          /* [n] sorted position -> original index */
    /* Morton-sorted particle state */
    double *pos4;           /* [n,4] x y z m              (32-byte aligned) */
    double *vel4;           /* [n,4] vx vy vz press/rho^2 (32-byte aligned) */
    float *rel4;            /* [n,4] fp32 position relative to the cell origin, cell code */
    /* neighbour structure */
    int32_t *nbr;           /* [ceil(n/32)*32 * max_nbrs] warp-transposed ELL rows */
    int32_t *cnt;           /* [n] neighbours per sorted particle */
    sph_status *status;     /* [1] */
    /* multi-GPU slab decomposition (all zero / NULL on one GPU) */
    const int32_t *n_valid; /* device scalar or NULL: particle slots >= *n_valid hold no particle (the ghost
                               region has a fixed capacity so that no count has to travel to the host); they
                               are binned into a spare cell behind the table that no pass visits */
    const int64_t *sort_key;/* [n] or NULL: ghosts (original index >= n_owned) are ordered by this key (their
                               global id) inside a cell instead of by their arrival slot, so results do not
                               depend on the order in which the neighbour rank packed them */
    int32_t n_owned;        /* > 0: original indices >= n_owned are ghosts -- binned and listed as neighbours of
                               owned particles, but no rows, densities or forces are computed FOR them.  On a grid
                               restricted by sph_grid_restrict_x the ghosts are the particles of the first and the
                               last local x layer; anything else raises SPH_F_OUT_OF_SLAB */
    int32_t reserved0;
    const uint32_t *group_tab; /* [sph_group_tab_elems(grid)] or NULL: what the neighbour pass needs to know about every
                               group of 2 x 2 x 2 cells and depends on the GRID only (the cell-code contributions of
                               the 4 + 4 + 4 cell layers around it, its base cell coordinates), written by
                               sph_group_table once per grid.  NULL: every block of sph_nlist_build works it out
                               again (~250 dependent instructions before its first load) */
} sph_buffers;

/* ------------------------------------------------------------------ host-side planning */

/* Choose the cell grid for a periodic box (replaces nothing in the reference, which scans
 * all n^2/2 pairs: neighbour_list.py:168-169).  occ_lo/occ_hi (may be NULL) are the extents
 * the particles occupy; dimensions that are mostly empty get coarser cells so that the cell
 * table stays O(n).  slab_lo/slab_layers (may be NULL) restrict the local grid to a range
 * of global x cell layers for the multi-GPU slab decomposition. */
int sph_grid_plan(const double box[3], double cutoff, double tolerance, int64_t n_hint,
                  const double *occ_lo, const double *occ_hi, sph_grid *grid);
int sph_grid_restrict_x(sph_grid *grid, int32_t first_layer, int32_t n_layers);

/* Elements needed in sph_buffers.scan_tmp for a grid with `ncode` cell codes. */
int64_t sph_scan_tmp_elems(uint32_t ncode);
/* Elements needed in sph_buffers.nbr. */
int64_t sph_nbr_elems(int32_t n, int32_t max_nbrs);
/* Elements (uint32) of sph_buffers.group_tab for this grid; 0 when the grid has no use for one. */
int64_t sph_group_tab_elems(const sph_grid *grid);
/* Fill buf->group_tab for `grid` (call again whenever the grid changes, sph_grid_restrict_x included). */
int sph_group_table(const sph_grid *grid, const sph_buffers *buf, void *stream);

/* ------------------------------------------------------------------ the hot path */

int sph_status_reset(sph_status *d_status, void *stream);

/* Cell list: counting-sort binning into block-Morton-ordered cells, deterministic order inside a
 * cell (ascending original index).  Fills code, rank, cell_count, cell_start, perm.
 * First half of VerletList.build (neighbour_list.py:160-189). */
int sph_cells_build(const sph_grid *grid, const sph_buffers *buf, const double *d_r, void *stream);

/* The same in three steps, for the slab decomposition: the owned particles are binned BEFORE the ghosts
 * arrive, and that pass also lists the particles of the two boundary cell layers (local x layers 1 and
 * ncl[0] - 2: sph_grid_restrict_x keeps one ghost layer on each side) -- the ghosts the x-neighbours need --
 * at no extra pass over the positions.
 *   begin   zeroes the cell counters and bins particles [first, first + count) (all of them hold particles:
 *           *buf->n_valid is not read here, it is written after this pass); d_idx_left / d_idx_right
 *           (each `cap` entries, may be NULL) receive the boundary-layer indices in arrival order, their
 *           numbers go to sph_status.halo_count (which may exceed cap: SPH_F_HALO_OVERFLOW at pack time)
 *   add     bins a further range (the ghost slots; slots >= *buf->n_valid go to the spare cell)
 *   finish  scan, scatter, canonical order inside the cells */
int sph_cells_begin(const sph_grid *grid, const sph_buffers *buf, const double *d_r, int32_t first, int32_t count,
                    int32_t *d_idx_left, int32_t *d_idx_right, int32_t cap, void *stream);
int sph_cells_add(const sph_grid *grid, const sph_buffers *buf, const double *d_r, int32_t first, int32_t count,
                  void *stream);
int sph_cells_finish(const sph_grid *grid, const sph_buffers *buf, void *stream);

/* Reorder particle state into the Morton-sorted working set (pos4, vel4, rel4) using
 * the existing perm.  Also what VerletList.separations amounts to when the list is kept
 * and only positions moved (neighbour_list.py:236-252): it refreshes the operands from
 * which every pair separation is recomputed on the fly. */
int sph_gather(const sph_grid *grid, const sph_buffers *buf, const double *d_r, const double *d_v,
               const double *d_m, void *stream);

/* Neighbour structure: for every particle the set {j : rsq_ij < thr} under the reference's
 * minimum image and predicate (neighbour_list.py:105-123,170-178), bit-exact, stored both
 * ways (j in row i and i in row j).  Second half of VerletList.build. */
int sph_nlist_build(const sph_grid *grid, const sph_buffers *buf, void *stream);

/* Density summation + van der Waals EOS: properties.spam_properties (properties.py:63-120),
 * f_properties.spam_properties (f_properties.py:16-147), c_properties.pyx:83-211.
 * Reads t and writes rho, p, pco, u, t in ORIGINAL particle order (d_t is in/out, as p.t
 * is in the reference) and leaves press/rho^2 in vel4[.,3] for sph_force.
 * `use_hlr` == 1: long-range density only (rho_lr with h = hlr, no EOS;
 * f_properties.py:102-107) -- d_p..d_t may then be NULL.
 * `use_hlr` == 2: density + EOS in SpamComplete's direction (spam_complete_force.py:134,151-152): d_u is
 * the integrated internal energy (read, not written), T = max((u + a rho) / kb, 0) is written to d_t and
 * the pressures follow from that T.
 * `list_fresh` != 0 promises positions are those the cell list was built from. */
int sph_density_eos(const sph_grid *grid, const sph_buffers *buf, const sph_eos *eos,
                    const double *d_h_orig, int h_uniform, int list_fresh, int use_hlr,
                    double *d_rho, double *d_p, double *d_pco, double *d_u, double *d_t,
                    void *stream);

/* Pair force and rates of change: forces.SpamForce.apply (forces.py:327-368),
 * SpamForce2d (:246-274) with dim == 2, c_forces.SpamForce.apply (c_forces.pyx:62-115);
 * with the cohesive pressure it is CohesiveSpamForce (forces.py:371-405,
 * c_forces.pyx:131-182).  d_press and d_rho (original order) give press_i / rho_i^2; pass
 * both NULL to reuse the values the preceding sph_density_eos left in vel4[.,3].
 * ACCUMULATES into vdot[n,3], udot[n] (original order) as the reference does; `first_force` != 0
 * says vdot / udot would be all zero at this point (particles.py:549-550 zeroes them before the first
 * force of an evaluation), so the results are stored instead and the caller need not zero them.
 * `part` splits the pass for the slab decomposition: 0 every particle, 1 all but the particles of the slab's two
 * boundary cell layers (local x layers 1 and ncl[0] - 2 of a restricted grid), 2 only those -- the only ones with
 * ghost neighbours, so part 1 can run while the ghosts' (p, rho) are still in flight. */
int sph_force(const sph_grid *grid, const sph_buffers *buf, const double *d_press,
              const double *d_rho, const double *d_h_orig, int h_uniform, int list_fresh,
              double fcutoff, int dim, int first_force, int part, double *d_vdot, double *d_udot,
              void *stream);

/* Refresh press_i / rho_i^2 (the operand sph_force gathers) from original-order arrays for the
 * particles whose original index is >= first_orig only.  The slab decomposition uses it for the
 * ghosts, whose (p, rho) arrive from the neighbouring rank after the density pass. */
int sph_pressure_term(const sph_buffers *buf, const double *d_press, const double *d_rho, int32_t first_orig,
                      void *stream);

/* Heat conduction from the full heat-flux vector: c_forces.SpamConduction.apply
 * (c_forces.pyx:196-239).  d_jq[n,3] and d_rho[n] are in original order; d_aux4 is caller-owned
 * scratch of n*4 doubles (32-byte aligned).  ACCUMULATES into udot[n] (original order). */
int sph_conduction(const sph_grid *grid, const sph_buffers *buf, const double *d_jq, const double *d_rho,
                   const double *d_h_orig, int h_uniform, int list_fresh, double *d_aux4, double *d_udot,
                   void *stream);

/* Velocity gradient and Newtonian viscous pair force: the eta / zeta terms of
 * spam_complete_force.SpamComplete (spam_complete_force.py:36-60,140-165).  BUILDER-DEFINED arithmetic:
 * the reference computes them in the Fortran routine sphforce3d, which it does not ship (SURVEY.md
 * section 8c), so these two follow the in-repo conventions instead --
 *   gradv_i[a][b] = sum_j (m_j / rho_j) (v_j - v_i)_a dW_ij/dx_b      (properties.py:95-98 with the
 *                   weight of c_properties.pyx:166-188 and the FINAL density, so the result does not
 *                   depend on the pair order as the reference's running-density loop does)
 *                   dW_ij is the gradient with respect to r_j - r_i as everywhere in the reference, so
 *                   gradv estimates MINUS the velocity gradient (v = A r gives gradv ~ -A)
 *   pi = 2 eta symmetric_traceless(gradv) + zeta tr(gradv) I             (tensor.py:5-16; this is
 *                   -2 eta S - zeta (div v) I for the true gradient: shear heats the fluid)
 *   a  = (pi_i / rho_i^2 + pi_j / rho_j^2) . dW_ij, +a to i, -a to j;  du = a . dv / 2,
 *   udot_i += du m_j, udot_j += du m_i                                  (forces.py:353-368, tensor for scalar)
 * d_gradv is [n,3,3] in original order; d_aux4 / d_aux8 are caller-owned scratch of n*4 / n*8 doubles
 * (32-byte aligned).  sph_viscous_force ACCUMULATES into vdot / udot like sph_force. */
int sph_gradv(const sph_grid *grid, const sph_buffers *buf, const double *d_rho, const double *d_h_orig,
              int h_uniform, int list_fresh, double *d_aux4, double *d_gradv, void *stream);
int sph_viscous_force(const sph_grid *grid, const sph_buffers *buf, const double *d_gradv, const double *d_rho,
                      double eta, double zeta, const double *d_h_orig, int h_uniform, int list_fresh,
                      double fcutoff, double *d_aux8, double *d_vdot, double *d_udot, void *stream);

/* The remaining terms of SpamComplete -- `cgrad` (density-gradient / capillary coefficient), `sigma`, `rcoef`
 * (repulsive core size and strength) and the heat-flux output `jq` (spam_complete_force.py:28-31,52-54,113-115,
 * 158-183).  BUILDER-DEFINED like the viscous term: their arithmetic is in the absent Fortran routine.
 *   sph_gradient      out_i = sum_j wgt_j (f_j - [subtract_self] f_i) grad_i W_ij, grad_i W_ij = -dW_ij/d(r_j - r_i);
 *                     d_f == NULL means f = 1.  (f = 1, wgt = m: density gradient; f = T, wgt = m / rho: grad T.)
 *   sph_stress_force  the pair force of sph_viscous_force for ANY symmetric stress tensor S[n,3,3] (original order):
 *                     a = (S_i / rho_i^2 + S_j / rho_j^2) . dW_ij.  SpamComplete feeds it the gradient part of the
 *                     Korteweg tensor, cgrad (g (x) g - |g|^2 I / 2) with g = grad rho_lr, on the long smoothing length.
 *   sph_core_force    phi(r) = rcoef (1 - r^2 / sigma^2)^4 per unit mass for r < sigma (Hoover's SPAM core):
 *                     a = -(8 rcoef / sigma^2) (1 - r^2 / sigma^2)^3 (r_j - r_i), +a to i, -a to j.
 * All three ACCUMULATE (forces) or overwrite (gradient) in original order; aux4 / aux8 as above. */
int sph_gradient(const sph_grid *grid, const sph_buffers *buf, const double *d_f, const double *d_wgt, int subtract_self,
                 const double *d_h_orig, int h_uniform, int list_fresh, double *d_aux4, double *d_out, void *stream);
int sph_stress_force(const sph_grid *grid, const sph_buffers *buf, const double *d_stress, const double *d_rho,
                     const double *d_h_orig, int h_uniform, int list_fresh, double fcutoff, double *d_aux8,
                     double *d_vdot, double *d_udot, void *stream);
int sph_core_force(const sph_grid *grid, const sph_buffers *buf, double sigma, double rcoef, int list_fresh,
                   double *d_vdot, double *d_udot, void *stream);

/* ------------------------------------------------------------------ pair-list API surface */

/* Lexicographic i<j pair list in ORIGINAL indices (what VerletList.build leaves in
 * nl.iap, neighbour_list.py:186-189).  Two calls: count fills d_row_count[n] (pairs per original i); the caller
 * scans it into d_row_start[n+1] as 64-BIT offsets (d_row_start[n] = nip: a 256 Mi-particle box holds 4.5e9 pairs,
 * more than 32 bits count); fill writes iap[nip,2] (int32 particle indices). */
int sph_pairs_count(const sph_buffers *buf, uint32_t *d_row_count, void *stream);
int sph_pairs_fill(const sph_buffers *buf, const int64_t *d_row_start, int32_t *d_iap,
                   int64_t cap_pairs, void *stream);
/* Exclusive 32-bit scan (used by sph_cells_build; exported for callers with short lists). */
int sph_exclusive_scan_u32(const uint32_t *d_in, uint32_t *d_out, uint32_t *d_tmp, int64_t n,
                           void *stream);

/* Per-pair separations for an explicit pair list: NeighbourList.separations
 * (neighbour_list.py:63-83) / pairsep.pairsep (pairsep.pyx:25-81) + minimum image. */
int sph_separations(const double box[3], const int32_t *d_iap, int64_t nip, const double *d_r,
                    const double *d_v, double *d_drij, double *d_rij, double *d_rsq,
                    double *d_dv, void *stream);

/* Per-pair kernel values: spkernel.lucy_kernel (spkernel.py:86-118) with h of the first
 * pair member (properties.py:88). */
int sph_pair_kernels(const int32_t *d_iap, int64_t nip, const double *d_rij, const double *d_drij,
                     const double *d_h, double *d_wij, double *d_dwij, void *stream);

/* VerletList.compress (neighbour_list.py:191-223): drop listed neighbours that fail the
 * predicate at the CURRENT sorted positions (after sph_gather). */
int sph_compress(const sph_grid *grid, const sph_buffers *buf, void *stream);

/* VerletList.ponder_rebuild (neighbour_list.py:225-234): status->rebuild = max|r_old-r|^2 > tol^2. */
int sph_ponder_rebuild(const double *d_r_old, const double *d_r, int32_t n, double tol_sq,
                       sph_status *d_status, void *stream);

/* ------------------------------------------------------------------ multi-GPU slab decomposition */

/* Particle arrays of one rank in pyticles' layout (owned particles in front, ghost slots behind). */
typedef struct sph_fields {
    double *r, *v;          /* [n,3] */
    double *m, *h, *t;      /* [n]   */
    int64_t *gid;           /* [n] global particle id */
} sph_fields;

#define SPH_HALO_COLS 10    /* doubles per halo row: r[3] v[3] m h t gid */

/* Ghost exchange A of the slab decomposition with fixed-capacity buffers, so that no
 * count ever travels to the host.  A send buffer is (cap + 1) rows of SPH_HALO_COLS doubles; row 0 is the
 * header {rows that follow, particles in the layer, 0...}.
 *   pack    rows of the boundary-layer particles d_idx_left / d_idx_right (from sph_cells_begin, counts in
 *           status->halo_count) into d_send_left / d_send_right; more than cap: SPH_F_HALO_OVERFLOW
 *   unpack  the buffers received from the left / right neighbour into the slots first, first + 1, ...
 *           (left neighbour's particles, then the right's); writes the number of valid slots
 *           first + ghosts to *d_n_valid and the two ghost counts to status->ghost_count */
int sph_halo_pack(const sph_fields *f, const int32_t *d_idx_left, const int32_t *d_idx_right, int32_t cap,
                  double *d_send_left, double *d_send_right, sph_status *d_status, void *stream);
int sph_halo_unpack(const sph_fields *f, const double *d_recv_left, const double *d_recv_right, int32_t cap,
                    int32_t first, int32_t *d_n_valid, sph_status *d_status, void *stream);
/* Ghost exchange B: two per-particle scalars (p and rho) of the same particles in the same order; buffers
 * are cap rows of 2 doubles, the counts are the ones exchange A left in the status block. */
int sph_halo_pack2(const int32_t *d_idx_left, const int32_t *d_idx_right, int32_t cap, const double *d_a,
                   const double *d_b, double *d_send_left, double *d_send_right, const sph_status *d_status,
                   void *stream);
int sph_halo_unpack2(const double *d_recv_left, const double *d_recv_right, int32_t cap, int32_t first,
                     double *d_a, double *d_b, const sph_status *d_status, void *stream);

/* ------------------------------------------------------------------ stepping helpers ("next" rows) */

/* x <- a + s * b over len doubles: the state-vector updates of integrator.py:37-41,44-59,62-95. */
int sph_axpy(double *d_x, const double *d_a, const double *d_b, double s, int64_t len, void *stream);

/* box.MirrorBox.apply (box.py:51-73) / box.PeriodicBox.apply (box.py:33-47). kind: 0 mirror, 1 periodic (the
 * reference's reset to the opposite face), 2 true periodic wrap x - L floor(x / L) (not in the reference) */
int sph_box_apply(const double box[3], int kind, double *d_r, double *d_v, int32_t n, void *stream);

const char *sph_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PYTICLES_B200_H */
