/*
 * turtlemd_b200 -- C ABI of the B200-native TurtleMD hot path.
 *
 * One shared library (libturtlemd_b200.so, built from turtlemd_b200/csrc/ for
 * sm_100a) exports everything below with C linkage.  Plain pointers and sizes
 * only; no torch / numpy types.  Every function returns a tmd_status
 * (0 = TMD_OK); tmd_last_error() gives the text of the last failure on the
 * calling thread.  There is NO CPU fallback: without a CUDA device every
 * compute entry point returns TMD_ERR_NO_DEVICE.
 *
 * Two families of entry points:
 *
 *   tmd_host_*   take HOST buffers laid out exactly like the reference's NumPy
 *                arrays (pos/vel/force (N,d) C-order float64, ptype (N,) int64,
 *                virial (d,d) float64).  They are what a reference maintainer
 *                binds behind Potential.force()/potential()/potential_and_force()
 *                and MDIntegrator.integration_step(): copy in, launch, copy out.
 *
 *   tmd_*        (no "host") take DEVICE pointers + a cudaStream_t (as void*),
 *                never synchronise and never allocate on the hot path (scratch
 *                comes from a per-device arena that only grows).  The Python
 *                package keeps the particle arrays resident in HBM and calls
 *                these; so do the resident runs (tmd_nlist_run_vv, tmd_slabrun_*).
 *
 * Reference = infretis/turtlemd 2023.3.1; citations are file:line in its tree.
 *
 * Array conventions
 *   pos, vel, force : double[n*dim], row-major (n, dim), dim in {1,2,3}
 *   imass           : double[n]            (reference: (n,1), system/particles.py:44)
 *   tidx            : int32[n]  dense type index 0..ntypes-1 (the host wrapper
 *                     maps the reference's arbitrary int64 ptype values onto it)
 *   table           : double[6*T*T] = lj1, lj2, lj3, lj4, offset, rcut2, each (T,T);
 *                     the values LennardJonesCut.set_parameters computes
 *                     (potentials/lennardjones.py:94-115)
 *   vw              : double[10] = { V, W[0][0..2], W[1][0..2], W[2][0..2] } ; the
 *                     (dim,dim) virial is the top-left block of the 3x3
 */
#ifndef TURTLEMD_B200_H
#define TURTLEMD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMD_ABI_VERSION 2

typedef enum tmd_status {
    TMD_OK = 0,
    TMD_ERR_INVALID = 1,     /* bad argument (dim, sizes, null pointer, unknown flag) */
    TMD_ERR_CUDA = 2,        /* a CUDA runtime call failed; see tmd_last_error()       */
    TMD_ERR_NO_DEVICE = 3,   /* no usable CUDA device: the product has no CPU path     */
    TMD_ERR_UNSUPPORTED = 4, /* valid in the reference but outside the hot path (e.g. ntypes > TMD_MAX_TYPES) */
    TMD_ERR_NOMEM = 5
} tmd_status;

/* what an evaluation should produce / how it combines with what is already there */
#define TMD_WANT_V 1u        /* potential energy  -> vw[0]                                      */
#define TMD_WANT_F 2u        /* forces            -> force[]                                    */
#define TMD_WANT_W 4u        /* virial            -> vw[1..9]                                   */
#define TMD_ACCUMULATE 8u    /* add to force[] / vw[] instead of overwriting (2nd, 3rd potential of a System) */
#define TMD_V_STANDALONE 16u /* LJ only: r6inv = 1/rsq^3 as LennardJonesCut.potential() does
                                (lennardjones.py:172) instead of (1/rsq)^3 (:254-255)           */

#define TMD_MAX_TYPES 8      /* particle types per LJ table on the device path */

typedef void *tmd_stream_t;  /* a cudaStream_t; NULL = the legacy default stream */

/* Orthorhombic box as Box/BoxBase store it (system/box.py:140-158).
 * Non-periodic directions: periodic=0; length/ilength are then ignored
 * (the reference keeps inf / 0 there).                                                         */
typedef struct tmd_box {
    int32_t dim;
    int32_t periodic[3];
    double length[3];
    double ilength[3];       /* the STORED reciprocal 1/length, multiplied with as box.py:402 does */
} tmd_box;

/* TriclinicBox (system/box.py:183-318): what pbc_dist needs, exactly as the host object stores it.     */
typedef struct tmd_cell {
    int32_t dim;              /* 2 or 3                                                             */
    int32_t reserved;
    double matrix[9];         /* box_matrix, row-major 3x3, zero padded (box.py:244-246)            */
    double inverse[9];        /* box_matrix_inv = np.linalg.inv(box_matrix) (:247)                  */
    double translations[81];  /* translations[g*3+k], g < 3^dim: box_matrix @ index (:248-256)      */
    double half_width;        /* shortest_width_half (:257-258)                                     */
} tmd_cell;

/* ---------------------------------------------------------------- library / device ---------- */
int tmd_abi_version(void);
const char *tmd_last_error(void);
const char *tmd_status_string(int status);
int tmd_device_count(int *count);
int tmd_set_device(int device);
/* name[<=256]; sm = 10*major+minor; bytes of HBM; number of SMs */
int tmd_device_info(int device, char *name, int *sm, size_t *total_mem, int *num_sms);
int tmd_stream_synchronize(tmd_stream_t stream);
/* kernels this library has launched so far in this process (all devices, all streams) */
int tmd_launch_count(int64_t *launches);

/* device memory + copies for bindings that do not bring their own allocator */
int tmd_malloc(void **dptr, size_t bytes);
int tmd_free(void *dptr);
int tmd_memset_async(void *dptr, int value, size_t bytes, tmd_stream_t stream);
int tmd_memcpy_h2d_async(void *dptr, const void *hptr, size_t bytes, tmd_stream_t stream);
int tmd_memcpy_d2h_async(void *hptr, const void *dptr, size_t bytes, tmd_stream_t stream);
int tmd_memcpy_d2d_async(void *dst, const void *src, size_t bytes, tmd_stream_t stream);
/* page-lock an existing host range (e.g. the NumPy arrays of a Particles object) */
int tmd_host_register(void *hptr, size_t bytes);
int tmd_host_unregister(void *hptr);
/* peer access for the atom-decomposed multi-GPU path (cudaIpc*, 64-byte handles) */
int tmd_ipc_get_handle(void *dptr, unsigned char handle[64]);
int tmd_ipc_open_handle(const unsigned char handle[64], void **dptr);
int tmd_ipc_close_handle(void *dptr);

/* ---------------------------------------------------------------- pair potentials ----------- */

/* K1. LennardJonesCut.potential_and_force / .force / .potential
 * (potentials/lennardjones.py:145-273) with Box.pbc_dist_matrix fused
 * (system/box.py:384-403): tiled all-pairs FP64, every one of the N(N-1)/2 pairs is
 * distance-checked.  Deterministic (no atomics).  first/count select the rows this call
 * computes forces for (atom decomposition: rank g passes its own slab and gets V and W for
 * that slab's share; first=0,count=n is the whole system).                                      */
int tmd_lj_allpairs(const double *pos, const int32_t *tidx, int64_t n, const tmd_box *box,
                    const double *table, int32_t ntypes, uint32_t flags,
                    int64_t first, int64_t count,
                    double *force, double *vw, tmd_stream_t stream);

/* K2+K3. Same contract, O(N): bins a wrapped copy of pos into cells of edge >= max rcut
 * (counting sort), then walks the 3^dim stencil.  The API-visible pos is left unwrapped, as
 * the reference integrators leave it (integrators.py:90).  Returns TMD_ERR_UNSUPPORTED when a
 * periodic direction holds fewer than 3 cells (then only tmd_lj_allpairs matches the
 * reference's single-nearest-image rule); tmd_lj_cells_usable() tells in advance.               */
int tmd_lj_cells_usable(int64_t n, const tmd_box *box, const double *table, int32_t ntypes, int *usable);
int tmd_lj_cells(const double *pos, const int32_t *tidx, int64_t n, const tmd_box *box,
                 const double *table, int32_t ntypes, uint32_t flags,
                 int64_t first, int64_t count,
                 double *force, double *vw, tmd_stream_t stream);

/* K2+K3 with a persistent Verlet neighbour list (the fast large-N path).  The handle keeps the
 * per-atom candidate lists (all pairs within rcut+skin at build time) and reuses them while no
 * atom has moved more than skin/2 since the build; the displacement check, the rebuild and the
 * traversal are all stream-ordered (no host round trip).  Forces are those of the full evaluation:
 * the list is an acceleration structure, not an approximation, and a row whose list outgrew its
 * capacity is evaluated from the cells instead (slow, still exact).  Same box requirements as
 * tmd_lj_cells with rcut+skin as the cell edge (tmd_lj_nlist_usable).  One handle per
 * (system, potential, device); not thread-safe.                                                  */
typedef struct tmd_nlist tmd_nlist;
int tmd_nlist_create(tmd_nlist **out, double skin);
int tmd_nlist_destroy(tmd_nlist *nl);
int tmd_lj_nlist_usable(int64_t n, const tmd_box *box, const double *table, int32_t ntypes, double skin,
                        int *usable);
/* builds done / calls made / rows that outgrew the list capacity (cumulative over builds; they
 * take the slow exact path).  Synchronises the device.                                           */
int tmd_nlist_stats(tmd_nlist *nl, int64_t *builds, int64_t *calls, int32_t *overflow);
int tmd_lj_nlist(tmd_nlist *nl, const double *pos, const int32_t *tidx, int64_t n, const tmd_box *box,
                 const double *table, int32_t ntypes, uint32_t flags,
                 int64_t first, int64_t count,
                 double *force, double *vw, tmd_stream_t stream);

/* Fused force + kick + drift: `steps` VelocityVerlet steps (integrators.py:84-94, the loop of
 * simulation/mdsimulation.py:69-90) over a LennardJonesCut force field in ONE call, resident on the
 * device.  pos / vel / force / imass are the device arrays of the reference's particles (atom order,
 * (n, dim) row-major; force = the forces at pos, as VelocityVerlet expects to find them).  The state is
 * moved into the list's cell order once; each step is then one list traversal whose epilogue applies the
 * closing half kick of the step and the opening half kick + drift of the next to the row atom, writes the
 * record the next traversal gathers and does the Verlet-list displacement test; rebuilds are decided on
 * the device (gated launches, no host read).  On return pos, vel, force and vw = (V, W) hold the state
 * after `steps` steps.  Roundings, list builds and summation order are those of
 * tmd_vv_kick_drift + tmd_lj_nlist + tmd_vv_kick on the same handle: the trajectory is bit-identical. */
int tmd_nlist_run_vv(tmd_nlist *nl, double *pos, double *vel, double *force, const double *imass,
                     const int32_t *tidx, int64_t n, const tmd_box *box, const double *table, int32_t ntypes,
                     double dt, int32_t steps, double *vw, tmd_stream_t stream);

/* Measurement hook for bench.py's roofline: while on, tmd_nlist_run_vv brackets every traversal launch with
 * CUDA events on its stream, waits for the run at the end (so: never inside a timed region) and accumulates
 * the kernel time; tmd_nlist_timing returns the sum in milliseconds and the number of launches in it. */
/* ... with a DoubleWellPair (potentials/well.py:230-286) as the system's second potential: the few atoms of the bonded
 * types (at most 128; idx0 / idx1 as for tmd_dwpair) are evaluated by one small kernel per step from the slot-ordered
 * state and their forces are added in the traversal's epilogue -- System.potential_and_force's sum (system/system.py:58-76),
 * bit-identical to the step-by-step path.                                                                              */
typedef struct tmd_dwpair_term {
    const int32_t *idx0;
    int32_t n0;
    const int32_t *idx1;
    int32_t n1, same_type;
    double rwidth, width2, height;
} tmd_dwpair_term;
int tmd_nlist_run_vv_dwpair(tmd_nlist *nl, double *pos, double *vel, double *force, const double *imass,
                            const int32_t *tidx, int64_t n, const tmd_box *box, const double *table, int32_t ntypes,
                            double dt, int32_t steps, const tmd_dwpair_term *term, double *vw, tmd_stream_t stream);
int tmd_nlist_set_timing(tmd_nlist *nl, int32_t on);
int tmd_nlist_timing(tmd_nlist *nl, double *traversal_ms, int64_t *traversals);

/* Slab path (multi-GPU; also the sorted single-GPU run).  The rows of a rank are the
 * cell-order SLOTS [slot_first, slot_first+slot_count) of the list's own ordering -- a spatially
 * compact slab -- and the rank's state lives in slot order: `records` (n x 4 doubles: x, y, z with
 * the lattice vector of the build removed, type bits) are what the pair kernel gathers from and
 * what the tmd_slab_* integrator kernels advance, so on several GPUs the records of the boundary
 * slots are the halo message.  There is no reference counterpart: the reference is one process
 * (potentials/lennardjones.py:226-273 is still the arithmetic being evaluated).
 *   tmd_nlist_bind_records  caller-owned record buffer (>= n*4 doubles, 32-byte aligned), before the first build
 *   tmd_nlist_build_slots   bin ALL n atoms of pos (atom order), lists for the given slot rows;
 *                           gated != 0: only if the handle's device flag [0] is raised
 *   tmd_nlist_force_slots   traversal of the current records; force_s is indexed by SLOT (n x dim)
 *   tmd_nlist_view          device pointers of the handle's tables (valid until the next reallocation) */
typedef struct tmd_nlist_buffers {
    double *records;        /* [n][4] current records, cell order                                      */
    double *build_records;  /* [n][4] records at build time (displacement test)                        */
    int32_t *order;         /* [n] slot -> atom                                                        */
    int32_t *slot_of;       /* [n] atom -> slot                                                        */
    double *image;          /* [n][3] by atom: lattice vector removed at build time                    */
    int32_t *flags;         /* [8]: [0] rebuild request, [1] rows on the slow path, [2] builds, [3] calls,
                               [4] / [5] slots below / above the row range that the lists reference,
                               [6] atoms in the fullest cell layer of the slowest direction              */
    int32_t cells[3];       /* cell grid                                                               */
    int32_t half_width;     /* stencil half width in cells                                             */
    int32_t capacity;       /* list entries per row                                                    */
} tmd_nlist_buffers;
int tmd_nlist_bind_records(tmd_nlist *nl, double *records);
int tmd_nlist_build_slots(tmd_nlist *nl, const double *pos, const int32_t *tidx, int64_t n, const tmd_box *box,
                          const double *table, int32_t ntypes, int64_t slot_first, int64_t slot_count,
                          int32_t gated, tmd_stream_t stream);
int tmd_nlist_force_slots(tmd_nlist *nl, uint32_t flags, double *force_s, double *vw, tmd_stream_t stream);
int tmd_nlist_view(tmd_nlist *nl, tmd_nlist_buffers *out);

/* ---- resident multi-GPU VelocityVerlet run: one spatial slab per rank ----------
 * K steps of integrators.py:84-94 over lennardjones.py:226-273 with the system cut into slabs of
 * cell layers along z, one per rank.  Per step and rank: ONE traversal whose epilogue integrates,
 * writes the next records and stores the two boundary layers into the neighbours' buffers (peer
 * memory), and one barrier kernel (mailboxes in peer memory) that also ORs the ranks' "rebuild"
 * flags; rebuilds (atom migration, ghost layers, binning, lists) are gated launches decided on
 * the device.  No host read and no collective library inside a step.  `window` is the only memory
 * other ranks touch: export it with tmd_ipc_get_handle, open the peers' handles, hand all of them
 * to tmd_slabrun_connect.  The barrier spins on what peers write: ranks must sit on DIFFERENT
 * GPUs; tests emulate ranks on one GPU phase by phase (tmd_slabrun_phase, signal and wait halves
 * apart).  Waits are bounded (~2 s): then err has bit 1 set and the run is void, the box lives.  */
typedef struct tmd_slabrun tmd_slabrun;
int tmd_slabrun_create(tmd_slabrun **out, int32_t world, int32_t rank, int64_t n_global, const tmd_box *box,
                       const double *table, int32_t ntypes, double skin, double dt);
int tmd_slabrun_destroy(tmd_slabrun *sr);
int tmd_slabrun_window(tmd_slabrun *sr, void **ptr, size_t *bytes);
int tmd_slabrun_connect(tmd_slabrun *sr, void *const *windows);
/* replicated DEVICE arrays of all n_global atoms (force valid for pos) -> slabs, ghosts, lists   */
int tmd_slabrun_load(tmd_slabrun *sr, const double *pos, const double *vel, const double *force, const double *imass,
                     const int32_t *tidx, tmd_stream_t stream);
/* a DoubleWellPair as the system's second potential (VelocityVerlet): HOST arrays of the global indices of the atoms of
 * types[0] / types[1] (ascending, at most 128 in all); their positions are replicated through the windows, every rank
 * evaluates the pairs of the bonded atoms it owns.  Call before tmd_slabrun_load.                     */
int tmd_slabrun_set_dwpair(tmd_slabrun *sr, const int32_t *gid0, int32_t n0, const int32_t *gid1, int32_t n1,
                           int32_t same_type, double rwidth, double width2, double height, const tmd_box *box);
int tmd_slabrun_start(tmd_slabrun *sr, tmd_stream_t stream);
int tmd_slabrun_step(tmd_slabrun *sr, int32_t last, tmd_stream_t stream);
/* start + steps - 1 full steps + one closing step; use_graph: pairs of steps replayed from a CUDA graph whose
 * rebuild launches sit in conditional IF nodes (a step without a rebuild is two kernels)             */
int tmd_slabrun_run(tmd_slabrun *sr, int32_t steps, int32_t use_graph, int64_t *graph_launches, tmd_stream_t stream);
/* one piece of a step (tests: several emulated ranks on one GPU, all signal halves of a barrier before any wait half) */
enum tmd_slabrun_phase_id {
    TMD_SR_MIGRATE = 0, TMD_SR_B0_SIGNAL = 1, TMD_SR_B0_WAIT = 2, TMD_SR_COLLECT = 3, TMD_SR_B1_SIGNAL = 4,
    TMD_SR_B1_WAIT = 5, TMD_SR_BUILD = 6, TMD_SR_B2_SIGNAL = 7, TMD_SR_B2_WAIT = 8, TMD_SR_TRAVERSE = 9,
    TMD_SR_B3_SIGNAL = 10, TMD_SR_B3_WAIT = 11, TMD_SR_PRE = 12, TMD_SR_ADVANCE = 13
};
int tmd_slabrun_phase(tmd_slabrun *sr, int32_t phase, int32_t last, tmd_stream_t stream);
/* own atoms -> their entries of the atom-ordered arrays (pos, vel, force all NULL: only vw[10], this rank's (V, W) share) */
int tmd_slabrun_export(tmd_slabrun *sr, double *pos, double *vel, double *force, double *vw, tmd_stream_t stream);
int tmd_slabrun_status(tmd_slabrun *sr, int32_t *err, int64_t *n_own, int64_t *builds, int32_t *layers);

/* K4. DoubleWell.potential / .force (potentials/well.py:65-81): V = sum a x^4 - b (x-c)^2 over
 * ALL n*dim coordinates, F = -4 a x^3 + 2 b (x - c), virial = 0.                                */
int tmd_doublewell(const double *pos, int64_t n, int32_t dim, double a, double b, double c,
                   uint32_t flags, double *force, double *vw, tmd_stream_t stream);

/* K5. DoubleWellPair.potential / .force (potentials/well.py:230-286), Box.pbc_dist fused
 * (system/box.py:370-382, strict '>').  No cut-off.  idx0[n0] / idx1[n1] are the (ascending)
 * indices of the atoms of types[0] / types[1]; same_type != 0 means types[0]==types[1] and
 * idx1 is ignored.  rwidth, width2, height are the derived parameters of set_parameters
 * (well.py:206-212).                                                                             */
int tmd_dwpair(const double *pos, int64_t n, const tmd_box *box,
               const int32_t *idx0, int32_t n0, const int32_t *idx1, int32_t n1, int32_t same_type,
               double rwidth, double width2, double height, uint32_t flags,
               double *force, double *vw, tmd_stream_t stream);

/* ---------------------------------------------------------------- integrators --------------- */
/* All element-wise over the flat n*dim array; arithmetic order follows the NumPy expressions
 * so that, given identical forces, the updates are bit-identical to the reference's.           */

/* VelocityVerlet (integrators.py:84-94): v += (dt/2 * F) * imass ; x += dt * v                  */
int tmd_vv_kick_drift(double *pos, double *vel, const double *force, const double *imass,
                      int64_t n, int32_t dim, double dt, tmd_stream_t stream);
/* ... second half kick: v += (dt/2 * F) * imass                                                  */
int tmd_vv_kick(double *vel, const double *force, const double *imass,
                int64_t n, int32_t dim, double dt, tmd_stream_t stream);
/* last kick of step k and first kick + drift of step k+1 in one pass (both roundings kept)     */
int tmd_vv_kick_kick_drift(double *pos, double *vel, const double *force, const double *imass,
                           int64_t n, int32_t dim, double dt, tmd_stream_t stream);

/* Verlet (integrators.py:53-68).  first_call != 0 initialises prev = x - v*dt first (:56-57). */
int tmd_verlet_step(double *pos, double *vel, double *prev, const double *force, const double *imass,
                    int64_t n, int32_t dim, double dt, int32_t first_call, tmd_stream_t stream);

/* Counter-based noise: "tmd-normal-v1" = Philox4x64-10 laid out like numpy.random.Philox(key)
 * + a fixed-arithmetic Box-Muller.  Draw number q of the stream is independent of
 * who asks for it, so any sharding of the particles reproduces the same trajectory.            */
typedef struct tmd_rng {
    uint64_t key;            /* Philox key word 0 (word 1 = 0), = numpy Philox(key=...)          */
    uint64_t offset;         /* index of this step's first normal in the stream                  */
    /* ABI 2.  Optional device-side step counter, honoured by tmd_langevin_inertia_pre / _post_pre:
     * when step_counter != NULL the kernel starts at offset + step_counter[0] * step_stride, so a
     * step recorded once in a CUDA graph draws fresh noise at every replay (tmd_counter_add
     * advances the counter inside the graph).  NULL: the host passes the position in `offset`.    */
    const uint64_t *step_counter;
    uint64_t step_stride;    /* normals one step of the WHOLE system consumes (2 n dim)           */
} tmd_rng;
/* counter[0] += increment on the stream (one thread)                                            */
int tmd_counter_add(uint64_t *counter, uint64_t increment, tmd_stream_t stream);

/* LangevinOverdamped (integrators.py:151-163), the part between force() and potential():
 *   xi = sigma_i * z[offset + (first+i)*dim + k] ;  x += bddt_i * F + xi ;  v = xi
 * sigma_i = sqrt(2 dt imass_i/(beta gamma)), bddt_i = dt*imass_i/gamma (:137-149) are computed
 * in-kernel.  noise != NULL: use these host-drawn normals (n*dim, already scaled by sigma,
 * i.e. the array rgen.normal returned) instead of the in-kernel stream.                         */
int tmd_langevin_overdamped(double *pos, double *vel, const double *force, const double *imass,
                            int64_t n, int32_t dim, double dt, double gamma, double beta,
                            const tmd_rng *rng, int64_t first, const double *noise,
                            tmd_stream_t stream);

/* Scalar constants of LangevinInertia.initiate_parameters (integrators.py:223-263).  The
 * per-particle ones follow in-kernel from imass_i:  a2_i = a2f*imass_i, b1_i = b1f*imass_i,
 * b2_i = b2f*imass_i, sig_r2 = ((dt*imass_i)/beta_gamma)*kfac, sig_v2 = (one_minus_e2*imass_i)/beta,
 * cov_rv = (imass_i/beta_gamma)*one_minus_e_sq and the 2x2 Cholesky factor L of
 * [[sig_r2, cov_rv], [cov_rv, sig_v2]] (:249-262).  The Python host fills this struct with NumPy
 * (same expressions as the reference); tmd_langevin_inertia_constants() is the libm version.    */
typedef struct tmd_langevin_consts {
    double c0;             /* exp(-gamma dt)                       */
    double a1;             /* c1*dt                                */
    double a2f;            /* c2*dt**2                             */
    double b1f;            /* (c1-c2)*dt                           */
    double b2f;            /* c2*dt                                */
    double dt;
    double beta;
    double beta_gamma;     /* beta*gamma                           */
    double kfac;           /* 2 - (3 - 4 e + e**2)/(gamma dt)      */
    double one_minus_e2;   /* 1 - e**2                             */
    double one_minus_e_sq; /* (1 - e)**2                           */
} tmd_langevin_consts;
int tmd_langevin_inertia_constants(double dt, double gamma, double beta, tmd_langevin_consts *out);
/* the resident slab run (tmd_slabrun_*) with LangevinInertia instead of VelocityVerlet: noise of the counter-based
 * stream (key, offset = position of the first step's first normal) addressed by GLOBAL atom index; call before the
 * first tmd_slabrun_start                                                                          */
int tmd_slabrun_set_langevin(tmd_slabrun *sr, const tmd_langevin_consts *consts, uint64_t key, uint64_t offset);
/* `steps` x LangevinInertia.integration_step (integrators.py:265-282) over LennardJonesCut without leaving the device: the
 * resident run of tmd_nlist_run_vv with the Langevin updates as ONE pass per step over the slot-ordered state (closing update
 * of a step + opening update of the next); noise = rng->key / rng->offset (position of the first step's first normal),
 * addressed by atom index; force() and potential() of a step come from one traversal (same numbers).  Bit-identical to
 * `steps` x (tmd_langevin_inertia_pre, tmd_lj_nlist, tmd_langevin_inertia_post).                                        */
int tmd_nlist_run_langevin(tmd_nlist *nl, double *pos, double *vel, double *force, const double *imass,
                           const int32_t *tidx, int64_t n, const tmd_box *box, const double *table, int32_t ntypes,
                           const tmd_langevin_consts *consts, const tmd_rng *rng, int32_t steps, double *vw,
                           tmd_stream_t stream);

/* LangevinInertia (integrators.py:265-282), part 1 (before force()):
 *   (n0, n1) = z[offset + (first+i)*2*dim + 2k + {0,1}]
 *   x += (a1 v + a2_i F) + L00_i n0 ;  v' = (c0 v + b1_i F) + (L10_i n0 + L11_i n1)
 * noise_pos/noise_vel != NULL: host-drawn pos_rand / vel_rand arrays instead (any duck-typed
 * rgen, e.g. numpy default_rng or the reference tests' FakeRandomGenerator).                    */
int tmd_langevin_inertia_pre(double *pos, double *vel, const double *force, const double *imass,
                             int64_t n, int32_t dim, const tmd_langevin_consts *consts,
                             const tmd_rng *rng, int64_t first,
                             const double *noise_pos, const double *noise_vel,
                             tmd_stream_t stream);
/* part 2 of one step and part 1 of the next in one pass over the state (both use the same F): what
 * LangevinInertia.run() issues between two force evaluations; in-kernel stream only.              */
int tmd_langevin_inertia_post_pre(double *pos, double *vel, const double *force, const double *imass,
                                  int64_t n, int32_t dim, const tmd_langevin_consts *consts,
                                  const tmd_rng *rng, int64_t first, tmd_stream_t stream);
/* The same two updates fed with STANDARD normals that are already on the device (tmd_npstream_normals below: the
 * reference's own numpy.random.Generator stream, integrators.py:156-160 and :312): overdamped element e takes
 * normals[e]; inertia element e takes normals[2e] and normals[2e+1] (multivariate_normal's reshape) and
 * chol[3i .. 3i+2] = (l00, l10, l11) of particle i, the reference's np.linalg.cholesky (integrators.py:261);
 * x-noise = l00 z0, v-noise = fma(l11, z1, l10 z0), the rounding of NumPy's norm @ cho.T.  post != 0: the previous
 * step's closing update v = v' + b2 F is applied first (what tmd_langevin_inertia_post_pre does).        */
int tmd_langevin_overdamped_normals(double *pos, double *vel, const double *force, const double *imass,
                                    int64_t n, int32_t dim, double dt, double gamma, double beta,
                                    const double *normals, tmd_stream_t stream);
int tmd_langevin_inertia_pre_normals(double *pos, double *vel, const double *force, const double *imass,
                                     int64_t n, int32_t dim, const tmd_langevin_consts *consts,
                                     const double *normals, const double *chol, int32_t post,
                                     tmd_stream_t stream);
/* ... part 2 (after force()): v = v' + b2_i * F                                                  */
int tmd_langevin_inertia_post(double *vel, const double *force, const double *imass,
                              int64_t n, int32_t dim, const tmd_langevin_consts *consts,
                              tmd_stream_t stream);

/* Fused replica kernel (config "1M independent DoubleWell Langevin replicas"): nsteps full
 * LangevinInertia steps of a system whose only potential is DoubleWell, state in registers,
 * force recomputed in-kernel.  On return force[] and vw[0] hold F and V of the final positions
 * exactly as nsteps reference steps would leave particles.force / particles.v_pot.  The rank
 * holds replicas first .. first+n-1 of n_global; the noise of replica i does not depend on the
 * sharding.                                                                                      */
int tmd_langevin_inertia_doublewell(double *pos, double *vel, double *force, const double *imass,
                                    int64_t n, int32_t dim, double a, double b, double c,
                                    const tmd_langevin_consts *consts,
                                    const tmd_rng *rng, int64_t first, int64_t n_global,
                                    int32_t nsteps, double *vw, tmd_stream_t stream);

/* fill out[count] with normals offset .. offset+count-1 of the stream (host-twin tests) */
int tmd_philox_normal_fill(double *out, int64_t count, const tmd_rng *rng, tmd_stream_t stream);
/* same numbers computed on the HOST by the same source (no device needed) */
int tmd_host_philox_normal_fill(double *out, int64_t count, uint64_t key, uint64_t offset);
int tmd_host_philox_raw_fill(uint64_t *out, int64_t count, uint64_t key, uint64_t offset);

/* ------------------------------------------------------------------ NumPy's normal stream on the device
 * numpy.random.Generator(PCG64).standard_normal(count) -- what the reference's rgen.normal(...) consumes
 * (turtlemd/integrators.py:156-160, :312; default_rng(seed) at :216; system/particles.py:260) -- generated ON THE
 * DEVICE, bit for bit: PCG64 (jump-ahead per chunk of 32 words) + NumPy's 256-layer ziggurat + glibc's log1p
 * restated for the tail (csrc/npy_normal.h, csrc/npy_normal.cu).  The generator state lives on the device; a
 * stock script keeps its trajectory without a host draw or an upload per step.
 *   seed:    state_inc = {state hi, state lo, inc hi, inc lo} of bit_generator.state (PCG64);
 *   normals: out[count] (device) <- the next `count` standard normals; enqueues four kernels, reads nothing back;
 *   state:   waits for the stream, returns the generator state {hi, lo}, the 64-bit words consumed since the
 *            seed and (tests) counters[2] = chunks walked again off the canonical chain by any pass, chunks placed off it; TMD_ERR_INVALID if a call's word
 *            window was too short (never observed: > 100 sigma) -- the normals of that call are then not valid;
 *   A stream belongs to the device that is current at create() and is used from one CUDA stream at a time.
 *   libm_variant: which glibc log1p build this process runs (1: FMA build, 0: baseline, -1: neither -- then
 *            create() refuses and callers draw on the host).                                              */
int tmd_npstream_libm_variant(void);
int tmd_npstream_create(void **handle);
int tmd_npstream_destroy(void *handle);
int tmd_npstream_seed(void *handle, const uint64_t *state_inc, tmd_stream_t stream);
int tmd_npstream_normals(void *handle, double *out, int64_t count, tmd_stream_t stream);
int tmd_npstream_state(void *handle, uint64_t *state, uint64_t *words_consumed, uint64_t *counters,
                       tmd_stream_t stream);
/* the same stream on the HOST by the same source (no device needed); state_inc is advanced in place */
int tmd_host_npstream_normals(uint64_t *state_inc, double *out, int64_t count, uint64_t *words_consumed);

/* ---------------------------------------------------------------- slab (slot-ordered) integrators */
/* The same updates on slot-ordered state (see the slab path above); rows = slots
 * [slot_first, slot_first+slot_count), arrays indexed by global slot.  The drift kernels also raise
 * *flag when a row has moved more than sqrt(half_skin2) from its build-time record.
 * VelocityVerlet first half (integrators.py:88-90); the second half is tmd_vv_kick on offset pointers. */
int tmd_slab_vv_kick_drift(double *records, double *vel_s, const double *force_s, const double *imass_s,
                           const double *build_records, int64_t slot_first, int64_t slot_count, int32_t dim,
                           double dt, double half_skin2, int32_t *flag, tmd_stream_t stream);
/* LangevinInertia part 1 (integrators.py:265-279); noise addressed by atom index order[slot], i.e.
 * the same draws as tmd_langevin_inertia_pre; part 2 is tmd_langevin_inertia_post on offset pointers. */
int tmd_slab_langevin_inertia_pre(double *records, double *vel_s, const double *force_s, const double *imass_s,
                                  const double *build_records, const int32_t *order, int64_t slot_first,
                                  int64_t slot_count, int32_t dim, const tmd_langevin_consts *consts,
                                  const tmd_rng *rng, double half_skin2, int32_t *flag, tmd_stream_t stream);
/* slot order -> atom order (pos = record + image); any of pos / vel / force may be NULL.
 * gate != NULL: do nothing unless *gate != 0 (device-side decision).                              */
int tmd_slab_unsort(const double *records, const double *vel_s, const double *force_s, const int32_t *order,
                    const double *image, int64_t n, int32_t dim, double *pos, double *vel, double *force,
                    const int32_t *gate, tmd_stream_t stream);
/* atom order -> slot order; any of vel_s / imass_s / force_s may be NULL                          */
int tmd_slab_sort(const int32_t *order, int64_t n, int32_t dim, const double *vel, const double *imass,
                  const double *force, double *vel_s, double *imass_s, double *force_s, const int32_t *gate,
                  tmd_stream_t stream);

/* ---------------------------------------------------------------- observables (next) */
/* kinetic energy tensor 0.5 sum m v(x)v (system/particles.py:156-168) -> kin[9] (3x3 block)   */
int tmd_kinetic_tensor(const double *vel, const double *mass, int64_t n, int32_t dim,
                       double *kin, tmd_stream_t stream);

/* ---------------------------------------------------------------- triclinic cells */
/* LennardJonesCut.{potential,force,potential_and_force} (potentials/lennardjones.py:145-273) with
 * box = TriclinicBox: minimum image by TriclinicBox.pbc_dist (system/box.py:271-296), applied in every
 * direction as the reference does.  Same flags / rows / outputs as tmd_lj_allpairs.                 */
int tmd_lj_triclinic(const double *pos, const int32_t *tidx, int64_t n, const tmd_cell *cell,
                     const double *table, int32_t ntypes, uint32_t flags, int64_t first, int64_t count,
                     double *force, double *vw, tmd_stream_t stream);
/* DoubleWellPair.potential / .force (potentials/well.py:230-286) with box = TriclinicBox; arguments as
 * tmd_dwpair / tmd_host_dwpair.                                                                       */
int tmd_dwpair_triclinic(const double *pos, int64_t n, const tmd_cell *cell, const int32_t *idx0, int32_t n0,
                         const int32_t *idx1, int32_t n1, int32_t same_type, double rwidth, double width2,
                         double height, uint32_t flags, double *force, double *vw, tmd_stream_t stream);
int tmd_host_dwpair_triclinic(const double *pos, const int64_t *ptype, int64_t n, const tmd_cell *cell,
                              int64_t type0, int64_t type1, double rwidth, double width2, double height,
                              uint32_t flags, double *force, double *vw);
/* the same with host buffers (ptype already mapped to 0..ntypes-1)                                 */
int tmd_host_lj_triclinic(const double *pos, const int64_t *tidx, int64_t n, const tmd_cell *cell,
                          const double *table, int32_t ntypes, uint32_t flags, double *force, double *vw);

/* ---------------------------------------------------------------- velocity set-up */
/* generate_maxwell_velocities step 1 (system/particles.py:260-262):
 *   vel[i,k] = sqrt(imass[i]) * z,  z = normal number  rng->offset + (first + i)*dim + k  of the stream
 * (the order in which rgen.normal(size=(n,dim)) hands them out).                                    */
int tmd_maxwell_draw(double *vel, const double *imass, int64_t n, int32_t dim, const tmd_rng *rng,
                     int64_t first, tmd_stream_t stream);
/* linear_momentum + mass.sum() (system/particles.py:135-137, :153):
 *   out[0] = sum_i mass[i],  out[1+k] = sum_i vel[i,k]*mass[i]   (device double[4], fixed summation order) */
int tmd_momentum(const double *vel, const double *mass, int64_t n, int32_t dim, double *out,
                 tmd_stream_t stream);
/* vel[i,k] -= shift[k] (zero_momentum, :153) when shift != NULL, then vel[i,k] *= scale (:269) when
 * do_scale != 0; shift is a HOST double[dim].  Two roundings, in the reference's order.             */
int tmd_vel_shift_scale(double *vel, int64_t n, int32_t dim, const double *shift, int32_t do_scale,
                        double scale, tmd_stream_t stream);

/* ---------------------------------------------------------------- host-buffer plug-in API ---- */
/* LennardJonesCut.{potential_and_force,force,potential}(system) with the reference's own
 * arrays: pos (n,dim) f64, ptype (n,) int64 already mapped to 0..ntypes-1, outputs force
 * (n,dim), vw[10].  method: 0 = choose (cells when usable and n >= 4096), 1 = all-pairs,
 * 2 = cells.  Synchronous.                                                                       */
int tmd_host_lj(const double *pos, const int64_t *tidx, int64_t n, const tmd_box *box,
                const double *table, int32_t ntypes, uint32_t flags,
                int32_t method, double *force, double *vw);
int tmd_host_doublewell(const double *pos, int64_t n, int32_t dim, double a, double b, double c,
                        uint32_t flags, double *force, double *vw);
int tmd_host_dwpair(const double *pos, const int64_t *ptype, int64_t n, const tmd_box *box,
                    int64_t type0, int64_t type1, double rwidth, double width2, double height,
                    uint32_t flags, double *force, double *vw);
/* One VelocityVerlet.integration_step(system) for a single-LJ-potential System, host arrays in
 * and out (pos, vel, force updated in place; vw[10] out).                                       */
int tmd_host_vv_lj_step(double *pos, double *vel, double *force, const double *imass,
                        const int64_t *tidx, int64_t n, const tmd_box *box,
                        const double *table, int32_t ntypes, int32_t method, double dt, double *vw);

/* ---------------------------------------------------------------- measurement helpers -------- */
/* FP64 pipe peak: DFMA-only kernel on every SM; returns achieved TFLOP/s (FMA = 2 flop).        */
int tmd_bench_dfma(int32_t repeats, double *tflops, double *milliseconds);
/* STREAM-style copy of `bytes` bytes; returns GB/s counting read + write.                       */
int tmd_bench_copy(size_t bytes, int32_t repeats, double *gbs, double *milliseconds);

#ifdef __cplusplus
}
#endif
#endif /* TURTLEMD_B200_H */
