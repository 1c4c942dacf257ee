// Resident multi-GPU VelocityVerlet run over a LennardJonesCut force field: one spatial SLAB per rank, no
// host decision and no collective library inside a step.  Included at
// the end of lj_nlist.cu: it reuses the binning, list and traversal kernels of the single-GPU resident run.
//
// Decomposition.  The cell grid of the neighbour list (edge >= rcut + skin) is cut along its slowest
// direction: rank r owns the cell layers [z0, z1), every rank the same number of them (the number of layers
// is rounded down to a multiple of the rank count: slightly thicker cells instead of uneven slabs).  A rank
// holds its own atoms plus GHOST copies of the two layers next to its slab, bins both over the GLOBAL grid
// (cells outside its layers stay empty) and builds lists for its own slots only.  Slots inside a cell ascend
// by global atom index on every rank, so a ghost layer is slot for slot the owner's layer: the halo is a
// contiguous block of records, and "my slot s of layer z" lives at "base of that layer there + (s - base
// here)" on the neighbour.
//
// One step on a rank (everything stream-ordered; launches that belong to a rebuild run only when the step's gate word is
// set -- the same on every rank, see barrier; inside the step graph they sit in a conditional IF node):
//     [migrate -> barrier -> collect + ghost send -> barrier -> bin -> lists -> state -> barrier]   (rebuild)
//     traversal whose epilogue does kick + kick + drift, writes the records of the next step and STORES the
//         records of the two boundary layers straight into the neighbours' buffers (peer memory over NVLink)
//     barrier: every rank writes (epoch, "one of my atoms moved more than skin / 2") into every rank's
//         mailbox and waits until all ranks of this epoch are there; the OR of the flags is the next gate.  On real
//         ranks the traversal block that finishes last runs it (one launch per step); emulated ranks and the
//         rebuild phases use sr_barrier_kernel.
// Variants: LangevinInertia (traversal leaves forces by slot, sr_langevin_kernel integrates and pushes the halo,
// barrier kernel); a DoubleWellPair as second potential (bonded atoms' positions replicated through the windows,
// sr_bond_force_kernel before and sr_bond_push_kernel after the traversal, barrier kernel).
// The barrier is a spin on memory the peers write: the ranks must run on different GPUs (or, for tests, one
// after the other on one GPU with the signal and wait halves launched separately: tmd_slabrun_phase).  Every
// wait is bounded (~2 s), then the run is marked failed instead of hanging the box.
//
// Reference semantics: integrators.py:84-94 over lennardjones.py:226-273, positions never wrapped.  The
// trajectory equals the single-GPU one up to summation order (the grid differs when the layer count is rounded).
#pragma once

namespace tmd {

constexpr int kSrMaxWorld = 16;
constexpr int kSrMigDoubles = 8;  // x y z vx vy vz 1/m (global index, type)
constexpr int kSrMaxBonded = 128;

enum SrErr {
    SR_OK = 0,
    SR_ERR_TIMEOUT = 1,        // a peer did not arrive at a barrier
    SR_ERR_FAR = 2,            // an atom left the slab by more than one layer between two rebuilds
    SR_ERR_MIG_FULL = 4,       // more migrants than the inbox holds
    SR_ERR_GHOST_FULL = 8,     // a boundary layer outgrew the ghost inbox
    SR_ERR_OWN_FULL = 16,      // more own atoms than capacity
    SR_ERR_LOC_FULL = 32,      // own + ghosts beyond capacity
    SR_ERR_LAYER = 64,         // a ghost layer does not match its owner's layer (populations differ)
};

struct SrWin {  // pointers into ONE rank's window (the only memory other ranks touch)
    double4 *rec[2];              // [cap_loc] records, double-buffered by step parity
    double *mig_in[2];            // [cap_mig][8] migrants from the rank below (0) / above (1)
    double4 *ghost_in[2];         // [cap_ghost] x y z (global index, type) of the layer below (0) / above (1)
    unsigned long long *mailbox;  // [2][world]: (epoch << 1 | flag) written by every rank
    int *mig_count;               // [2]
    int *ghost_count;             // [2]
    int *info;                    // [4]: slot where my ghost layer below / above starts, their populations
    int *err;
    double *active[2];            // [kSrMaxBonded][3]: unwrapped positions of the atoms of a DoubleWellPair's types, written
                                  // by their owners into EVERY rank's table, double-buffered like the records
};

struct SrCounters {  // device words of a rank (private)
    int stay, out[2], ghost[2];  // appended by atomics during a rebuild; zeroed by the step barrier
    int n_own;
    int pad[2];
};

__device__ __forceinline__ unsigned long long pack_gid_type(int gid, int type) {
    return ((unsigned long long)(unsigned)type << 32) | (unsigned)gid;
}

__device__ __forceinline__ int sr_layer(double z, const NlGrid &g) {
    const double w = z - floor(z * g.ilen[2]) * g.len[2];  // exactly nl_bin_kernel's wrap
    int c = (int)(w * g.scale[2]);
    return c < 0 ? 0 : (c >= g.nc[2] ? g.nc[2] - 1 : c);
}

struct SrArgs {
    NlGrid grid;
    int world, rank, z0, z1, zbelow, zabove, layer;  // own layers [z0, z1); ghost layers zbelow, zabove; cells per layer
    int64_t cap_own, cap_loc, cap_mig, cap_ghost;
    SrWin self, below, above;
    SrCounters *cnt;
    int *dyn;              // [0] atoms binned (own + ghosts)
    const int *gate;       // NULL = always
    int *halo;             // see NlArgs::halo
    // staging (local index order): own atoms first, then ghosts
    double *pos_w, *vel_w, *imass_w, *force_w;
    int *gid_w, *type_w;
    // slot-ordered state of the own slots
    double *xu, *vs, *fs, *image_s, *imass_s;
    int *gid_s;
    // list structures
    const int *starts, *order;
    const double4 *xb;
    const double *image;
    int with_force;
    double *fx;            // NULL, or slot-ordered forces of the second potential (zeroed at every rebuild)
    int bond_n;            // atoms of a DoubleWellPair's types (0 = none)
    const int *bond_gid;
    int *bond_slot;
};

// ---- initial distribution from the replicated arrays (every rank holds the whole system at the start) ----
// mode 0: own atoms -> staging [0, n_own); mode 1: atoms of the two ghost layers -> staging [n_own, ...)
__global__ void __launch_bounds__(256) sr_load_kernel(const SrArgs s, int64_t n_global, const double *__restrict__ pos,
                                                      const double *__restrict__ vel, const double *__restrict__ force,
                                                      const double *__restrict__ imass, const int32_t *__restrict__ tidx,
                                                      int mode) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_global) return;
    const int zc = sr_layer(pos[i * 3 + 2], s.grid);
    if (mode == 0) {
        if (zc < s.z0 || zc >= s.z1) return;
        const int k = atomicAdd(&s.cnt->stay, 1);
        if (k >= s.cap_own) { atomicOr(s.self.err, SR_ERR_OWN_FULL); return; }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            s.pos_w[(int64_t)k * 3 + c] = pos[i * 3 + c];
            s.vel_w[(int64_t)k * 3 + c] = vel[i * 3 + c];
            s.force_w[(int64_t)k * 3 + c] = force[i * 3 + c];
        }
        s.imass_w[k] = imass[i];
        s.gid_w[k] = (int)i;
        s.type_w[k] = tidx ? tidx[i] : 0;
    } else {
        if (zc != s.zbelow && zc != s.zabove) return;
        const int k = s.cnt->stay + atomicAdd(&s.cnt->ghost[0], 1);
        if (k >= s.cap_loc) { atomicOr(s.self.err, SR_ERR_LOC_FULL); return; }
#pragma unroll
        for (int c = 0; c < 3; ++c) s.pos_w[(int64_t)k * 3 + c] = pos[i * 3 + c];
        s.gid_w[k] = (int)i;
        s.type_w[k] = tidx ? tidx[i] : 0;
    }
}

__global__ void sr_load_done_kernel(const SrArgs s) {
    s.cnt->n_own = s.cnt->stay;
    s.dyn[0] = s.cnt->stay + s.cnt->ghost[0];
}

// ---- rebuild, phase 1: every own atom goes to the staging area of the rank that owns its new layer ----
__global__ void __launch_bounds__(256) sr_migrate_kernel(const SrArgs s) {
    if (s.gate && *s.gate == 0) return;
    const int first = s.starts[s.z0 * s.layer], count = s.starts[s.z1 * s.layer] - first;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int slot = first + t;
    double x[3], v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        x[c] = s.xu[(int64_t)slot * 3 + c];
        v[c] = s.vs[(int64_t)slot * 3 + c];
    }
    const double im = s.imass_s[slot];
    const int gid = s.gid_s[slot];
    const int type = (int)__double_as_longlong(s.xb[slot].w);
    const int zc = sr_layer(x[2], s.grid);
    if (zc >= s.z0 && zc < s.z1) {
        const int k = atomicAdd(&s.cnt->stay, 1);  // <= the old population: fits
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            s.pos_w[(int64_t)k * 3 + c] = x[c];
            s.vel_w[(int64_t)k * 3 + c] = v[c];
        }
        s.imass_w[k] = im;
        s.gid_w[k] = gid;
        s.type_w[k] = type;
        return;
    }
    const int dir = zc == s.zbelow ? 0 : (zc == s.zabove ? 1 : -1);
    if (dir < 0) { atomicOr(s.self.err, SR_ERR_FAR); return; }
    const int k = atomicAdd(&s.cnt->out[dir], 1);
    if (k >= s.cap_mig) { atomicOr(s.self.err, SR_ERR_MIG_FULL); return; }
    // the rank below receives it "from above" (inbox 1), the rank above "from below" (inbox 0)
    double *dst = (dir == 0 ? s.below.mig_in[1] : s.above.mig_in[0]) + (int64_t)k * kSrMigDoubles;
    dst[0] = x[0]; dst[1] = x[1]; dst[2] = x[2];
    dst[3] = v[0]; dst[4] = v[1]; dst[5] = v[2];
    dst[6] = im;
    dst[7] = __longlong_as_double((long long)pack_gid_type(gid, type));
}

// ---- rebuild, phase 2: immigrants join the staging area; atoms of the two boundary layers go to the
//      neighbours' ghost inboxes ----
__global__ void __launch_bounds__(256) sr_collect_kernel(const SrArgs s) {
    if (s.gate && *s.gate == 0) return;
    const int n_stay = s.cnt->stay;
    const int in0 = min(s.self.mig_count[0], (int)s.cap_mig), in1 = min(s.self.mig_count[1], (int)s.cap_mig);
    const int total = n_stay + in0 + in1;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        s.cnt->n_own = min(total, (int)s.cap_own);
        if (total > s.cap_own) atomicOr(s.self.err, SR_ERR_OWN_FULL);
    }
    if (t >= total || t >= s.cap_own) return;
    double x[3];
    int gid, type;
    if (t >= n_stay) {
        const int q = t - n_stay;
        const double *src = (q < in0 ? s.self.mig_in[0] + (int64_t)q * kSrMigDoubles
                                     : s.self.mig_in[1] + (int64_t)(q - in0) * kSrMigDoubles);
        const unsigned long long pk = (unsigned long long)__double_as_longlong(src[7]);
        gid = (int)(unsigned)pk;
        type = (int)(pk >> 32);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            x[c] = src[c];
            s.pos_w[(int64_t)t * 3 + c] = x[c];
            s.vel_w[(int64_t)t * 3 + c] = src[3 + c];
        }
        s.imass_w[t] = src[6];
        s.gid_w[t] = gid;
        s.type_w[t] = type;
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) x[c] = s.pos_w[(int64_t)t * 3 + c];
        gid = s.gid_w[t];
        type = s.type_w[t];
    }
    const int zc = sr_layer(x[2], s.grid);
    const double4 g4 = make_double4(x[0], x[1], x[2], __longlong_as_double((long long)pack_gid_type(gid, type)));
    if (zc == s.z0) {  // my lowest layer is the ghost layer ABOVE the slab of the rank below
        const int k = atomicAdd(&s.cnt->ghost[0], 1);
        if (k < s.cap_ghost) s.below.ghost_in[1][k] = g4;
        else atomicOr(s.self.err, SR_ERR_GHOST_FULL);
    }
    if (zc == s.z1 - 1) {
        const int k = atomicAdd(&s.cnt->ghost[1], 1);
        if (k < s.cap_ghost) s.above.ghost_in[0][k] = g4;
        else atomicOr(s.self.err, SR_ERR_GHOST_FULL);
    }
}

// ---- rebuild, phase 3a: ghosts join the staging area behind the own atoms ----
__global__ void __launch_bounds__(256) sr_ghost_kernel(const SrArgs s) {
    if (s.gate && *s.gate == 0) return;
    const int n_own = s.cnt->n_own;
    const int g0 = min(s.self.ghost_count[0], (int)s.cap_ghost), g1 = min(s.self.ghost_count[1], (int)s.cap_ghost);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        s.dyn[0] = min(n_own + g0 + g1, (int)s.cap_loc);
        if (n_own + g0 + g1 > s.cap_loc) atomicOr(s.self.err, SR_ERR_LOC_FULL);
    }
    if (t >= g0 + g1 || n_own + t >= s.cap_loc) return;
    const double4 g4 = t < g0 ? s.self.ghost_in[0][t] : s.self.ghost_in[1][t - g0];
    const unsigned long long pk = (unsigned long long)__double_as_longlong(g4.w);
    const int64_t k = n_own + t;
    s.pos_w[k * 3 + 0] = g4.x;
    s.pos_w[k * 3 + 1] = g4.y;
    s.pos_w[k * 3 + 2] = g4.z;
    s.gid_w[k] = (int)(unsigned)pk;
    s.type_w[k] = (int)(pk >> 32);
}

// ---- rebuild, phase 3c (after binning and lists): the state of the own atoms moves to their new slots; what
//      the neighbours need to address my ghost layers is published ----
__global__ void __launch_bounds__(256) sr_state_kernel(const SrArgs s) {
    if (s.gate && *s.gate == 0) return;
    const int first = s.starts[s.z0 * s.layer], count = s.starts[s.z1 * s.layer] - first;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        s.self.info[0] = s.starts[s.zbelow * s.layer];
        s.self.info[1] = s.starts[s.zabove * s.layer];
        s.self.info[2] = s.starts[(s.zbelow + 1) * s.layer] - s.starts[s.zbelow * s.layer];
        s.self.info[3] = s.starts[(s.zabove + 1) * s.layer] - s.starts[s.zabove * s.layer];
        s.halo[0] = first;
        s.halo[1] = s.starts[(s.z0 + 1) * s.layer];
        s.halo[3] = s.starts[(s.z1 - 1) * s.layer];
        s.halo[4] = first + count;
        if (count != s.cnt->n_own) atomicOr(s.self.err, SR_ERR_LAYER);  // every own atom must sit in an own layer
    }
    if (t >= count) return;
    const int slot = first + t;
    const int64_t atom = s.order[slot];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        s.xu[(int64_t)slot * 3 + c] = s.pos_w[atom * 3 + c];
        s.vs[(int64_t)slot * 3 + c] = s.vel_w[atom * 3 + c];
        s.image_s[(int64_t)slot * 3 + c] = s.image[atom * 3 + c];
        if (s.with_force) s.fs[(int64_t)slot * 3 + c] = s.force_w[atom * 3 + c];
    }
    s.imass_s[slot] = s.imass_w[atom];
    const int g = s.gid_w[atom];
    s.gid_s[slot] = g;
    if (s.fx) {
#pragma unroll
        for (int c = 0; c < 3; ++c) s.fx[(int64_t)slot * 3 + c] = 0.0;
        for (int k = 0; k < s.bond_n; ++k)
            if (s.bond_gid[k] == g) s.bond_slot[k] = slot;  // (an atom of both lists of a two-type bond cannot exist)
    }
}

// ---- a DoubleWellPair (well.py:230-286) as the system's second potential: a handful of atoms, no cut-off ----
// Their unwrapped positions are replicated: the owner of a bonded atom stores it into every rank's table after each position
// update (before the step barrier); every rank then evaluates the pairs of the bonded atoms it OWNS into a slot-ordered
// force array that the traversal's epilogue adds (System.potential_and_force's sum, system.py:58-76).  Arithmetic and
// partner order are dwpair_kernel's (wells.cu).
struct SrBond {
    int n0, n1, same_type, nown;
    double rwidth, width2, height;
    double len[3], ilen[3], half[3];
    int per[3];
    const int *gid;    // [nown]: global indices of the atoms of types[0], then of types[1] (each ascending)
    int *my_slot;      // [nown]: the own slot of that atom, -1 when another rank owns it (set at every rebuild)
    double *partials;  // [10]: this rank's share of the pairs' V and W (every pair seen from both ends)
};

__global__ void sr_bond_reset_kernel(const SrBond b, const int *gate) {
    if (gate && *gate == 0) return;
    if (threadIdx.x < b.nown) b.my_slot[threadIdx.x] = -1;
}

struct SrTables {
    double *tab[kSrMaxWorld];
};

// owners -> every rank's table of the given parity
__global__ void sr_bond_push_kernel(const SrBond b, const double *__restrict__ xu, SrTables t, int world) {
    const int k = threadIdx.x;
    if (k >= b.nown) return;
    const int slot = b.my_slot[k];
    if (slot < 0) return;
    const double x = xu[(int64_t)slot * 3 + 0], y = xu[(int64_t)slot * 3 + 1], z = xu[(int64_t)slot * 3 + 2];
    for (int r = 0; r < world; ++r) {
        t.tab[r][k * 3 + 0] = x;
        t.tab[r][k * 3 + 1] = y;
        t.tab[r][k * 3 + 2] = z;
    }
}

__global__ void __launch_bounds__(128) sr_bond_force_kernel(const SrBond b, const double *__restrict__ table,
                                                            double *__restrict__ fx, uint32_t flags) {
    __shared__ double smem[10 * 4];
    double acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.0;
    const int t = threadIdx.x;
    const int slot = t < b.nown ? b.my_slot[t] : -1;
    if (slot >= 0) {
        const bool in0 = t < b.n0;
        const int me = b.gid[t];
        const int obase = b.same_type ? 0 : (in0 ? b.n0 : 0);
        const int nother = b.same_type ? b.n0 : (in0 ? b.n1 : b.n0);
        double xi[3], fi[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < 3; ++k) xi[k] = table[t * 3 + k];
        const double height4 = __dmul_rn(4.0, b.height);  // well.py:212
        for (int p = 0; p < nother; ++p) {
            const int q = obase + p;
            const int you = b.gid[q];
            if (you == me) continue;
            const bool first = me < you;  // the reference visits the pair as (i, j) with i < j (particles.py:129-132)
            double d[3];
            double rsq = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double xj = table[q * 3 + k];
                double dk = first ? __dsub_rn(xi[k], xj) : __dsub_rn(xj, xi[k]);
                if (b.per[k] && fabs(dk) > b.half[k])  // Box.pbc_dist: strict '>' half box, stored reciprocal (box.py:376-378)
                    dk = __dsub_rn(dk, __dmul_rn(rint(__dmul_rn(dk, b.ilen[k])), b.len[k]));
                d[k] = dk;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) rsq = (k == 0) ? __dmul_rn(d[k], d[k]) : __dadd_rn(rsq, __dmul_rn(d[k], d[k]));
            const double delr = __dsqrt_rn(rsq);
            const double diff = __dsub_rn(delr, b.rwidth);
            const double qq = __ddiv_rn(__dmul_rn(diff, diff), b.width2);
            if (flags & TMD_WANT_V) {
                const double om = __dsub_rn(1.0, qq);
                acc[0] += __dmul_rn(b.height, __dmul_rn(om, om));
            }
            const double mag = __dmul_rn(__dmul_rn(height4, __dsub_rn(1.0, qq)), __ddiv_rn(diff, b.width2));
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double fk = __ddiv_rn(__dmul_rn(mag, d[k]), delr);
                fi[k] += first ? fk : -fk;
#pragma unroll
                for (int m = 0; m < 3; ++m) acc[1 + k * 3 + m] += __dmul_rn(fk, d[m]);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) fx[(int64_t)slot * 3 + k] = fi[k];
    }
    block_sum<10, 128>(acc, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) b.partials[k] = acc[k];
    }
}

// ---- barrier across the ranks + the few words that travel with it ----
struct SrBarrier {
    int world, rank;
    const int *gate;                          // NULL = always
    int mode;                                 // bit 0: signal, bit 1: wait
    int kind;                                 // 0: migrant counts, 1: ghost counts, 2: halo addresses, 3: step (flag OR)
    unsigned long long *mail[kSrMaxWorld];    // every rank's mailbox [2][world]
    unsigned long long *epoch;                // this rank's barrier count (device-resident: CUDA graph replays advance it)
    SrWin self, below, above;
    SrCounters *cnt;
    int *halo;
    int *local_flag;                          // kind 3: raised by this rank's epilogues
    int *gate_out;                            // kind 3: the OR over all ranks, the gate of the next step's rebuild
    int *builds;                              // nlflags of the list handle ([2] builds)
    int *work;                                // kind 3: the traversal's row counter, reset for the next launch
    unsigned long long *noise_steps;          // kind 3, Langevin: steps whose noise has been drawn (advance != 0: + 1)
    int noise_advance;
    cudaGraphConditionalHandle cond;          // kind 3 inside a CUDA graph: the IF node around the next step's rebuild
    int has_cond;
};

__global__ void sr_barrier_kernel(const SrBarrier b) {
    if (b.gate && *b.gate == 0) return;
    const int lane = threadIdx.x;
    unsigned long long e = *b.epoch;
    if (b.mode & 1) {
        e += 1;
        int flag = 0;
        if (lane == 0) {
            if (b.kind == 0) {  // the rank below reads my count of atoms moving down as "from above", and vice versa
                b.below.mig_count[1] = b.cnt->out[0];
                b.above.mig_count[0] = b.cnt->out[1];
            } else if (b.kind == 1) {
                b.below.ghost_count[1] = b.cnt->ghost[0];
                b.above.ghost_count[0] = b.cnt->ghost[1];
            }
        }
        if (b.kind == 3) flag = *b.local_flag != 0;
        __syncwarp();
        __threadfence_system();  // everything this rank wrote before (earlier kernels included) is ahead of the signal
        if (lane < b.world) {
            volatile unsigned long long *dst = b.mail[lane] + (e & 1ull) * b.world + b.rank;
            *dst = (e << 1) | (unsigned long long)flag;
        }
        __syncwarp();
        if (lane == 0) *b.epoch = e;
    }
    if (b.mode & 2) {
        int flag = 0, failed = 0;
        if (lane < b.world) {
            volatile unsigned long long *src = b.mail[b.rank] + (e & 1ull) * b.world + lane;
            unsigned long long t0, now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            unsigned long long v = *src;
            while ((v >> 1) != e) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (now - t0 > 2000000000ull) { failed = 1; break; }
                __nanosleep(64);
                v = *src;
            }
            flag = (int)(v & 1ull);
        }
        failed = __any_sync(0xffffffffu, failed);
        flag = __any_sync(0xffffffffu, flag);
        __threadfence_system();
        if (lane == 0) {
            if (failed) atomicOr(b.self.err, SR_ERR_TIMEOUT);
            if (b.kind == 2) {  // where my two boundary layers live on the neighbours
                b.halo[2] = b.below.info[1];
                b.halo[5] = b.above.info[0];
                if (b.below.info[3] != b.halo[1] - b.halo[0] || b.above.info[2] != b.halo[4] - b.halo[3])
                    atomicOr(b.self.err, SR_ERR_LAYER);
            } else if (b.kind == 3) {
                if (b.has_cond) cudaGraphSetConditional(b.cond, flag ? 1u : 0u);
                *b.gate_out = flag;
                *b.local_flag = 0;
                *b.work = 0;
                if (b.noise_advance) *b.noise_steps += 1;
                b.cnt->stay = 0;
                b.cnt->out[0] = b.cnt->out[1] = 0;
                b.cnt->ghost[0] = b.cnt->ghost[1] = 0;
            }
        }
    }
}

// first node of a step graph: the IF node around the first step's rebuild takes the gate the previous launch left
__global__ void sr_cond_kernel(cudaGraphConditionalHandle cond, const int *gate) {
    cudaGraphSetConditional(cond, *gate != 0 ? 1u : 0u);
}

// ---- before the first step: opening half kick + drift (integrators.py:88-90), the records the first traversal
//      reads, the halo copies and the displacement test ----
__global__ void __launch_bounds__(256) sr_pre_kernel(const SrArgs s, double4 *__restrict__ rec, double4 *peer_lo,
                                                     double4 *peer_hi, double dt, double half, double half_skin2,
                                                     int *flag) {
    const int first = s.halo[0], count = s.halo[4] - first;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int slot = first + t;
    const double im = s.imass_s[slot];
    const double4 b = s.xb[slot];
    double r[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double v = kick(s.vs[(int64_t)slot * 3 + c], s.fs[(int64_t)slot * 3 + c], im, half);
        const double x = __dadd_rn(s.xu[(int64_t)slot * 3 + c], __dmul_rn(dt, v));
        s.vs[(int64_t)slot * 3 + c] = v;
        s.xu[(int64_t)slot * 3 + c] = x;
        r[c] = __dsub_rn(x, s.image_s[(int64_t)slot * 3 + c]);
    }
    const double4 out = make_double4(r[0], r[1], r[2], b.w);
    rec[slot] = out;
    if (slot < s.halo[1]) peer_lo[s.halo[2] + (slot - first)] = out;
    if (slot >= s.halo[3]) peer_hi[s.halo[5] + (slot - s.halo[3])] = out;
    double d2 = (r[0] - b.x) * (r[0] - b.x);
    d2 = fma(r[1] - b.y, r[1] - b.y, d2);
    d2 = fma(r[2] - b.z, r[2] - b.z, d2);
    if (!(d2 <= half_skin2)) *flag = 1;
}

// LangevinInertia on the own slots (integrators.py:265-282).  mode bit 0: the closing update v = v' + b2 F of the step
// whose forces were just computed (:281); bit 1: the opening update of the next step (:270-279) with the noise of a
// coordinate addressed by its GLOBAL atom index (draws 2 (atom dim + k) and + 1 of the step, as the reference consumes
// them: any decomposition retraces the single-GPU trajectory), the records of the next traversal, the halo copies and
// the displacement test.  The step's position in the stream is offset + step_counter[0] * stride (device-resident).
__global__ void __launch_bounds__(256) sr_langevin_kernel(const SrArgs s, double4 *__restrict__ rec, double4 *peer_lo,
                                                          double4 *peer_hi, tmd_langevin_consts c, uint64_t key,
                                                          uint64_t offset, const unsigned long long *step_counter,
                                                          uint64_t stride, int mode, double half_skin2, int *flag) {
    const int first = s.halo[0], count = s.halo[4] - first;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int slot = first + t;
    const double im = s.imass_s[slot];
    const InertiaPerParticle p = inertia_constants(c, im);
    const uint64_t q_base = offset + (uint64_t)step_counter[0] * stride;
    const int64_t atom = s.gid_s[slot];
    const double4 b = s.xb[slot];
    double r[3];
    double z6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // even stream position: the six normals of this atom from two Philox blocks
    const uint64_t q0 = q_base + 2ull * (uint64_t)(atom * 3);
    const bool six = (mode & 2) && (q0 & 1ull) == 0;
    if (six) six_normals(key, q0, z6);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double f = s.fs[(int64_t)slot * 3 + k];
        double v = s.vs[(int64_t)slot * 3 + k];
        double x = s.xu[(int64_t)slot * 3 + k];
        if (mode & 1) v = __dadd_rn(v, __dmul_rn(__dmul_rn(c.b2f, im), f));  // exactly inertia_post_kernel
        if (mode & 2) {
            double n0, n1;
            if (six) { n0 = z6[2 * k]; n1 = z6[2 * k + 1]; }
            else two_normals(key, q_base + 2ull * (uint64_t)(atom * 3 + k), true, n0, n1);
            const double xr = __dmul_rn(p.l00, n0);  // norm @ cho.T  (integrators.py:315)
            const double vr = __dadd_rn(__dmul_rn(p.l10, n0), __dmul_rn(p.l11, n1));
            inertia_pre_update(x, v, f, xr, vr, c, p);
            s.xu[(int64_t)slot * 3 + k] = x;
        }
        s.vs[(int64_t)slot * 3 + k] = v;
        r[k] = __dsub_rn(x, s.image_s[(int64_t)slot * 3 + k]);
    }
    if (!(mode & 2)) return;
    const double4 out = make_double4(r[0], r[1], r[2], b.w);
    rec[slot] = out;
    if (slot < s.halo[1]) peer_lo[s.halo[2] + (slot - first)] = out;
    if (slot >= s.halo[3]) peer_hi[s.halo[5] + (slot - s.halo[3])] = out;
    double d2 = (r[0] - b.x) * (r[0] - b.x);
    d2 = fma(r[1] - b.y, r[1] - b.y, d2);
    d2 = fma(r[2] - b.z, r[2] - b.z, d2);
    if (!(d2 <= half_skin2)) *flag = 1;
}

// own atoms -> the entries of the global, atom-ordered arrays (the caller sums / gathers over the ranks)
__global__ void __launch_bounds__(256) sr_export_kernel(const SrArgs s, double *__restrict__ pos, double *__restrict__ vel,
                                                        double *__restrict__ force) {
    const int first = s.halo[0], count = s.halo[4] - first;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int slot = first + t;
    const int64_t g = s.gid_s[slot];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        pos[g * 3 + c] = s.xu[(int64_t)slot * 3 + c];
        vel[g * 3 + c] = s.vs[(int64_t)slot * 3 + c];
        force[g * 3 + c] = s.fs[(int64_t)slot * 3 + c];
    }
}

}  // namespace tmd

struct tmd_slabrun {
    int world, rank, device, ntypes;
    int64_t n_global;
    double skin, dt;
    // device memory
    void *window;
    size_t window_bytes;
    void *priv;  // one private allocation, carved up below
    size_t priv_bytes;
    tmd::SrArgs s;
    tmd::NlArgs a;
    tmd::SrBarrier bar;
    double4 *rec_self[2], *rec_below[2], *rec_above[2];
    int *flags;           // nlflags[16] | dyn[4] | gates[2] | local_flag[2] | halo[8]
    int *gates, *local_flag, *work;
    tmd::NlStepBarrier *step_bar;  // device copy: the traversal's last block runs the step barrier on real ranks
    // a DoubleWellPair as second potential (tmd_slabrun_set_dwpair)
    int bonded;
    tmd::SrBond bond;
    double *fx;
    double *active_self[2];
    tmd::SrTables active_all[2];
    // LangevinInertia instead of VelocityVerlet (tmd_slabrun_set_langevin)
    int langevin;
    tmd_langevin_consts lconsts;
    uint64_t noise_key, noise_offset, noise_stride;
    unsigned long long *noise_steps;
    double *partial_vw, *vw;
    int fblocks, bin_blocks, list_blocks;
    int connected, loaded;
    int64_t step;         // host-side: steps enqueued so far (buffer and gate parity)
    int p;                // record buffer the next traversal reads
    // two steps as one CUDA graph: the rebuild of each sits in an IF node decided on the device
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    cudaStream_t capture_stream[2];
    int graph_failed;
};

namespace tmd {

static size_t sr_align(size_t x) { return (x + 255) & ~(size_t)255; }

// carve a rank's window (same layout on every rank: computed from the shared capacities)
static SrWin sr_window_layout(void *base, int world, int64_t cap_loc, int64_t cap_mig, int64_t cap_ghost, size_t *bytes) {
    char *p = reinterpret_cast<char *>(base);
    size_t off = 0;
    SrWin w;
    auto take = [&](size_t n) { char *q = p + off; off += sr_align(n); return q; };
    for (int k = 0; k < 2; ++k) w.rec[k] = reinterpret_cast<double4 *>(take((size_t)cap_loc * sizeof(double4)));
    for (int k = 0; k < 2; ++k) w.mig_in[k] = reinterpret_cast<double *>(take((size_t)cap_mig * kSrMigDoubles * sizeof(double)));
    for (int k = 0; k < 2; ++k) w.ghost_in[k] = reinterpret_cast<double4 *>(take((size_t)cap_ghost * sizeof(double4)));
    w.mailbox = reinterpret_cast<unsigned long long *>(take((size_t)2 * world * sizeof(unsigned long long)));
    w.mig_count = reinterpret_cast<int *>(take(2 * sizeof(int)));
    w.ghost_count = reinterpret_cast<int *>(take(2 * sizeof(int)));
    w.info = reinterpret_cast<int *>(take(4 * sizeof(int)));
    w.err = reinterpret_cast<int *>(take(sizeof(int)));
    for (int k = 0; k < 2; ++k) w.active[k] = reinterpret_cast<double *>(take((size_t)kSrMaxBonded * 3 * sizeof(double)));
    if (bytes) *bytes = off;
    return w;
}

// the slab grid: as nl_make_grid, but the number of layers along the slowest direction is a multiple of `world`
static bool sr_make_grid(const tmd_box *box, double reach, int64_t n, int world, NlGrid &g) {
    if (box->dim != 3) return false;
    if (!nl_make_grid(box, reach, n, g)) return false;
    if (g.h != 1) return false;
    const int nz = (g.nc[2] / world) * world;
    if (nz < 2 * world || nz < 3) return false;  // at least two layers per rank: own and ghost layers never coincide
    g.nc[2] = nz;
    g.scale[2] = nz / g.len[2];
    g.ncells = g.nc[0] * g.nc[1] * g.nc[2];
    return true;
}

static void sr_launch_barrier(tmd_slabrun *sr, int kind, int mode, const int *gate, cudaStream_t st,
                              const cudaGraphConditionalHandle *cond = nullptr, int noise_advance = 0) {
    SrBarrier b = sr->bar;
    b.noise_advance = noise_advance;
    b.kind = kind;
    b.mode = mode;
    b.gate = gate;
    b.has_cond = cond != nullptr;
    if (cond) b.cond = *cond;
    if (kind == 3) {
        const int g = (int)(sr->step & 1);
        b.local_flag = sr->local_flag;
        b.gate_out = sr->gates + (g ^ 1);
    }
    ::tmd::count_launch(), sr_barrier_kernel<<<1, 32, 0, st>>>(b);
}

}  // namespace tmd

using namespace tmd;

extern "C" {

int tmd_slabrun_destroy(tmd_slabrun *sr) {
    if (!sr) return TMD_OK;
    if (sr->exec) cudaGraphExecDestroy(sr->exec);
    if (sr->graph) cudaGraphDestroy(sr->graph);
    for (int k = 0; k < 2; ++k)
        if (sr->capture_stream[k]) cudaStreamDestroy(sr->capture_stream[k]);
    cudaFree(sr->window);
    cudaFree(sr->priv);
    delete sr;
    return TMD_OK;
}

int tmd_slabrun_create(tmd_slabrun **out, int32_t world, int32_t rank, int64_t n_global, const tmd_box *box,
                       const double *table, int32_t ntypes, double skin, double dt) {
    TMD_REQUIRE(out != nullptr, "out is NULL");
    TMD_REQUIRE(world >= 2 && world <= kSrMaxWorld && rank >= 0 && rank < world, "needs 2 <= world <= 16 and 0 <= rank < world");
    TMD_REQUIRE(box && box->dim == 3, "slab runs need a three-dimensional box");
    TMD_REQUIRE(n_global >= 2 && skin > 0.0 && dt > 0.0, "needs n >= 2, skin > 0, dt > 0");
    TMD_CHECK(require_device());
    tmd_slabrun *sr = new (std::nothrow) tmd_slabrun();
    if (!sr) return TMD_ERR_NOMEM;
    memset(sr, 0, sizeof(*sr));
    sr->world = world;
    sr->rank = rank;
    sr->n_global = n_global;
    sr->ntypes = ntypes;
    sr->skin = skin;
    sr->dt = dt;
    TMD_CUDA(cudaGetDevice(&sr->device));
    NlArgs &a = sr->a;
    memset(&a, 0, sizeof(a));
    double rc2max = 0.0;
    int status = load_lj_table(table, ntypes, a.tab, rc2max);
    if (status != TMD_OK) { delete sr; return status; }
    a.uniform_table = 1;
    for (int c = 0; c < 6; ++c)
        for (int p = 1; p < ntypes * ntypes; ++p)
            if (memcmp(&a.tab.c[c][p], &a.tab.c[c][0], sizeof(double)) != 0) a.uniform_table = 0;
    const double reach = sqrt(rc2max) + skin;
    if (!sr_make_grid(box, reach, n_global, world, a.grid)) {
        delete sr;
        set_error("tmd_slabrun: needs a fully periodic 3-D box with at least %d cell layers of edge rcut+skin along z "
                  "(two per rank); use fewer ranks or the atom decomposition", 2 * world);
        return TMD_ERR_UNSUPPORTED;
    }
    const NlGrid &g = a.grid;
    const int per = g.nc[2] / world;
    SrArgs &s = sr->s;
    memset(&s, 0, sizeof(s));
    s.grid = g;
    s.world = world;
    s.rank = rank;
    s.z0 = rank * per;
    s.z1 = s.z0 + per;
    s.zbelow = (s.z0 - 1 + g.nc[2]) % g.nc[2];
    s.zabove = s.z1 % g.nc[2];
    s.layer = g.nc[0] * g.nc[1];
    // capacities: density fluctuations of a fluid are small next to these margins; a run that outgrows them
    // stops with an error code, never silently
    const double mean_layer = (double)n_global / g.nc[2];
    s.cap_ghost = (int64_t)(mean_layer * 1.5) + 1024;
    s.cap_own = (int64_t)(mean_layer * per * 1.25) + 2048;
    s.cap_loc = s.cap_own + 2 * s.cap_ghost;
    s.cap_mig = (int64_t)(mean_layer * 0.25) + 1024;
    if (s.cap_loc >= (1ll << 27)) {
        delete sr;
        set_error("tmd_slabrun: too many atoms per rank");
        return TMD_ERR_UNSUPPORTED;
    }

    // ---- the window: what the other ranks write ----
    sr_window_layout(nullptr, world, s.cap_loc, s.cap_mig, s.cap_ghost, &sr->window_bytes);
    if (cudaMalloc(&sr->window, sr->window_bytes) != cudaSuccess) { cudaGetLastError(); delete sr; return TMD_ERR_NOMEM; }
    cudaMemset(sr->window, 0, sr->window_bytes);
    s.self = sr_window_layout(sr->window, world, s.cap_loc, s.cap_mig, s.cap_ghost, nullptr);

    // ---- private memory ----
    double volume = box->length[0] * box->length[1] * box->length[2];
    int cap = (int)(1.35 * (double)n_global / volume * 4.18879020478639 * reach * reach * reach) + 24;
    cap = (cap + kBlockEntries - 1) & ~(kBlockEntries - 1);
    if (cap > 1024) cap = 1024;
    const int64_t L = s.cap_loc, O = s.cap_own;
    const size_t ncells1 = (size_t)g.ncells + 1;
    const size_t chunks = ((size_t)g.ncells + kScanChunk - 1) / kScanChunk;
    sr->fblocks = nl_force_blocks((n_global + world - 1) / world, 3);
    if (nl_env_int("TMD_SR_FBLOCKS", 0) > 0) sr->fblocks = nl_env_int("TMD_SR_FBLOCKS", 0);  // tuning runs
    size_t off = 0;
    auto reserve = [&](size_t n) { const size_t o = off; off += sr_align(n); return o; };
    const size_t o_flags = reserve(64 * sizeof(int));
    const size_t o_cnt = reserve(sizeof(SrCounters));
    const size_t o_epoch = reserve(sizeof(unsigned long long));
    const size_t o_stepbar = reserve(sizeof(NlStepBarrier));
    const size_t o_noise = reserve(sizeof(unsigned long long));
    const size_t o_fx = reserve((size_t)L * 3 * sizeof(double));
    const size_t o_bond = reserve((size_t)kSrMaxBonded * 2 * sizeof(int) + 16 * sizeof(double));
    const size_t o_ints = reserve((2 * ncells1 + chunks + 5 * (size_t)L) * sizeof(int));
    const size_t o_wrapped = reserve((size_t)L * sizeof(double4));
    const size_t o_xb = reserve((size_t)L * sizeof(double4));
    const size_t o_xbf = reserve(((size_t)(L + 1) / 2 + kChunk / 2) * 2 * sizeof(float4));
    const size_t o_image = reserve((size_t)L * 3 * sizeof(double));
    const size_t o_entries = reserve((size_t)O * cap * sizeof(unsigned));
    const size_t o_nneigh = reserve((size_t)O * sizeof(int));
    const size_t o_posw = reserve((size_t)L * 3 * sizeof(double));
    const size_t o_velw = reserve((size_t)O * 3 * sizeof(double));
    const size_t o_forcew = reserve((size_t)O * 3 * sizeof(double));
    const size_t o_imassw = reserve((size_t)O * sizeof(double));
    const size_t o_gidw = reserve((size_t)L * sizeof(int));
    const size_t o_typew = reserve((size_t)L * sizeof(int));
    const size_t o_md = reserve((size_t)L * 13 * sizeof(double));
    const size_t o_gids = reserve((size_t)L * sizeof(int));
    const size_t o_pvw = reserve((size_t)sr->fblocks * 10 * sizeof(double));
    const size_t o_vw = reserve(16 * sizeof(double));
    sr->priv_bytes = off;
    if (cudaMalloc(&sr->priv, off) != cudaSuccess) { cudaGetLastError(); cudaFree(sr->window); delete sr; return TMD_ERR_NOMEM; }
    cudaMemset(sr->priv, 0, off);
    char *base = reinterpret_cast<char *>(sr->priv);
    sr->flags = reinterpret_cast<int *>(base + o_flags);
    sr->gates = sr->flags + 20;
    sr->local_flag = sr->flags + 22;
    sr->work = sr->flags + 32;
    sr->step_bar = reinterpret_cast<NlStepBarrier *>(base + o_stepbar);
    sr->noise_steps = reinterpret_cast<unsigned long long *>(base + o_noise);
    sr->noise_stride = 2ull * (uint64_t)n_global * 3ull;
    sr->fx = reinterpret_cast<double *>(base + o_fx);
    sr->bond.gid = reinterpret_cast<int *>(base + o_bond);
    sr->bond.my_slot = reinterpret_cast<int *>(base + o_bond) + kSrMaxBonded;
    sr->bond.partials = reinterpret_cast<double *>(base + o_bond + (size_t)kSrMaxBonded * 2 * sizeof(int));
    s.dyn = sr->flags + 16;
    s.halo = sr->flags + 24;
    s.cnt = reinterpret_cast<SrCounters *>(base + o_cnt);
    int *ints = reinterpret_cast<int *>(base + o_ints);
    a.counts = ints;
    a.starts = ints + ncells1;
    a.chunk_sums = ints + 2 * ncells1;
    a.cell_of = ints + 2 * ncells1 + chunks;
    a.rank_of = a.cell_of + L;
    a.order = a.cell_of + 2 * L;
    a.slot_of = a.cell_of + 3 * L;
    a.order_tmp = a.cell_of + 4 * L;
    a.wrapped = reinterpret_cast<double4 *>(base + o_wrapped);
    a.xb = reinterpret_cast<double4 *>(base + o_xb);
    a.xbf = reinterpret_cast<float4 *>(base + o_xbf);
    a.image = reinterpret_cast<double *>(base + o_image);
    a.entries = reinterpret_cast<unsigned *>(base + o_entries);
    a.nneigh = reinterpret_cast<int *>(base + o_nneigh);
    a.cap = cap;
    s.pos_w = reinterpret_cast<double *>(base + o_posw);
    s.vel_w = reinterpret_cast<double *>(base + o_velw);
    s.force_w = reinterpret_cast<double *>(base + o_forcew);
    s.imass_w = reinterpret_cast<double *>(base + o_imassw);
    s.gid_w = reinterpret_cast<int *>(base + o_gidw);
    s.type_w = reinterpret_cast<int *>(base + o_typew);
    double *md = reinterpret_cast<double *>(base + o_md);
    s.xu = md;
    s.vs = md + 3 * L;
    s.fs = md + 6 * L;
    s.image_s = md + 9 * L;
    s.imass_s = md + 12 * L;
    s.gid_s = reinterpret_cast<int *>(base + o_gids);
    s.starts = a.starts;
    s.order = a.order;
    s.xb = a.xb;
    s.image = a.image;
    sr->partial_vw = reinterpret_cast<double *>(base + o_pvw);
    sr->vw = reinterpret_cast<double *>(base + o_vw);

    // ---- the list / traversal arguments that never change ----
    a.pos = s.pos_w;
    a.tidx = s.type_w;
    a.n = L;  // capacity; the kernels read the population from dyn
    a.first = 0;
    a.count = O;
    a.row_mode = 1;
    a.ntypes = ntypes;
    a.flags = TMD_WANT_V | TMD_WANT_F | TMD_WANT_W;
    a.nlflags = sr->flags;
    a.dyn = s.dyn;
    a.row_cell_lo = s.z0 * s.layer;
    a.row_cell_hi = s.z1 * s.layer;
    a.gid = s.gid_w;
    a.halo = s.halo;
    a.work = sr->work;
    // measured at eight ranks (profiles/r02x_ab_g8.txt): claiming rows dynamically costs 1 % instead of saving the tail
    a.dynamic_rows = nl_env_int("TMD_SR_DYNAMIC", 0) != 0;
    {
        double lmax = fmax(box->length[0], fmax(box->length[1], box->length[2]));
        const double rf = reach + lmax * (1.0 / 1048576.0);
        a.reach2f = (float)(rf * rf * (1.0 + 1e-6));
    }
    a.half_skin2 = 0.25 * skin * skin;
    for (int k = 0; k < 3; ++k) a.centre[k] = 0.5 * box->length[k];
    a.force = nullptr;
    a.partial_vw = sr->partial_vw;
    a.vw = sr->vw;
    a.xu = s.xu;
    a.vs = s.vs;
    a.fs = s.fs;
    a.fx = nullptr;
    a.imass_s = s.imass_s;
    a.image_s = s.image_s;
    a.dt = dt;
    a.half_dt = dt * 0.5;  // VelocityVerlet.half_timestep (integrators.py:82)

    {
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nl_bin_kernel<3>, kBinThreads, 0);
        sr->bin_blocks = (per_sm < 1 ? 1 : per_sm) * num_sms();
        per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nl_list_kernel<3>, kNlThreads, 0);
        sr->list_blocks = (per_sm < 1 ? 1 : per_sm) * num_sms();
    }
    SrBarrier &b = sr->bar;
    memset(&b, 0, sizeof(b));
    b.world = world;
    b.rank = rank;
    b.epoch = reinterpret_cast<unsigned long long *>(base + o_epoch);
    b.self = s.self;
    b.cnt = s.cnt;
    b.halo = s.halo;
    b.builds = sr->flags;
    b.work = sr->work;
    b.noise_steps = sr->noise_steps;
    *out = sr;
    return TMD_OK;
}

int tmd_slabrun_window(tmd_slabrun *sr, void **ptr, size_t *bytes) {
    TMD_REQUIRE(sr && ptr && bytes, "NULL argument");
    *ptr = sr->window;
    *bytes = sr->window_bytes;
    return TMD_OK;
}

/* windows[r] = rank r's window as a device pointer valid in THIS process (tmd_ipc_open_handle of the handle rank r
 * exported; the own entry is the own window) */
int tmd_slabrun_connect(tmd_slabrun *sr, void *const *windows) {
    TMD_REQUIRE(sr && windows, "NULL argument");
    const int w = sr->world, below = (sr->rank - 1 + w) % w, above = (sr->rank + 1) % w;
    for (int r = 0; r < w; ++r) TMD_REQUIRE(windows[r] != nullptr, "a window pointer is NULL");
    TMD_REQUIRE(windows[sr->rank] == sr->window, "windows[rank] must be this rank's own window");
    SrArgs &s = sr->s;
    s.below = sr_window_layout(windows[below], w, s.cap_loc, s.cap_mig, s.cap_ghost, nullptr);
    s.above = sr_window_layout(windows[above], w, s.cap_loc, s.cap_mig, s.cap_ghost, nullptr);
    sr->bar.below = s.below;
    sr->bar.above = s.above;
    for (int r = 0; r < w; ++r)
        sr->bar.mail[r] = sr_window_layout(windows[r], w, s.cap_loc, s.cap_mig, s.cap_ghost, nullptr).mailbox;
    for (int k = 0; k < 2; ++k) {
        sr->rec_self[k] = s.self.rec[k];
        sr->rec_below[k] = s.below.rec[k];
        sr->rec_above[k] = s.above.rec[k];
    }
    for (int k = 0; k < 2; ++k) {
        sr->active_self[k] = s.self.active[k];
        for (int r = 0; r < w; ++r)
            sr->active_all[k].tab[r] = sr_window_layout(windows[r], w, s.cap_loc, s.cap_mig, s.cap_ghost, nullptr).active[k];
    }
    NlStepBarrier nb;
    memset(&nb, 0, sizeof(nb));
    nb.world = w;
    nb.rank = sr->rank;
    for (int r = 0; r < w; ++r) nb.mail[r] = sr->bar.mail[r];
    nb.epoch = sr->bar.epoch;
    nb.local_flag = sr->local_flag;
    nb.gates = sr->gates;
    nb.counters = reinterpret_cast<int *>(s.cnt);
    nb.err = s.self.err;
    TMD_CUDA(cudaMemcpy(sr->step_bar, &nb, sizeof(nb), cudaMemcpyHostToDevice));
    sr->connected = 1;
    return TMD_OK;
}

static int sr_build(tmd_slabrun *sr, const int *gate, int with_force, cudaStream_t st) {
    SrArgs s = sr->s;
    s.gate = gate;
    s.with_force = with_force;
    NlArgs a = sr->a;
    a.gate = gate ? gate : sr->flags + 30;  // flags[30] is kept at 1: "always"
    a.xs = sr->rec_self[sr->p];
    const int gb = (int)((s.cap_ghost * 2 + 255) / 256), ob = (int)((s.cap_own + 255) / 256);
    if (!with_force)  // (load: the ghosts are already staged)
        ::tmd::count_launch(), sr_ghost_kernel<<<gb, 256, 0, st>>>(s);
    void *args[] = {(void *)&a};
    ::tmd::count_launch();
    TMD_CUDA(cudaLaunchCooperativeKernel((const void *)nl_bin_kernel<3>, dim3(sr->bin_blocks), dim3(kBinThreads), args, 0, st));
    ::tmd::count_launch(), nl_list_kernel<3><<<sr->list_blocks, kNlThreads, 0, st>>>(a);
    if (sr->bonded) {
        s.fx = sr->fx;
        s.bond_n = sr->bond.nown;
        s.bond_gid = sr->bond.gid;
        s.bond_slot = sr->bond.my_slot;
        ::tmd::count_launch(), sr_bond_reset_kernel<<<1, kSrMaxBonded, 0, st>>>(sr->bond, gate);
    }
    ::tmd::count_launch(), sr_state_kernel<<<ob, 256, 0, st>>>(s);
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

/* Distribute the replicated system (device arrays of n_global atoms, forces valid for pos) and build the lists.
 * Follow with tmd_slabrun_start on every rank. */
int tmd_slabrun_load(tmd_slabrun *sr, const double *pos, const double *vel, const double *force, const double *imass,
                     const int32_t *tidx, tmd_stream_t stream) {
    TMD_REQUIRE(sr && pos && vel && force && imass, "NULL argument");
    TMD_REQUIRE(sr->connected, "call tmd_slabrun_connect first");
    TMD_REQUIRE(sr->ntypes == 1 || tidx, "tidx is required when ntypes > 1");
    cudaStream_t st = as_stream(stream);
    TMD_CUDA(cudaMemsetAsync(sr->flags, 0, 64 * sizeof(int), st));
    TMD_CUDA(cudaMemsetAsync(sr->flags + 30, 1, 1, st));  // the "always" gate word (non-zero)
    TMD_CUDA(cudaMemsetAsync(sr->s.cnt, 0, sizeof(SrCounters), st));
    const int nb = (int)((sr->n_global + 255) / 256);
    for (int mode = 0; mode < 2; ++mode)
        ::tmd::count_launch(), sr_load_kernel<<<nb, 256, 0, st>>>(sr->s, sr->n_global, pos, vel, force, imass, tidx, mode);
    ::tmd::count_launch(), sr_load_done_kernel<<<1, 1, 0, st>>>(sr->s);
    sr->step = 0;
    sr->p = 0;
    TMD_CHECK(sr_build(sr, nullptr, 1, st));
    sr->loaded = 1;
    return TMD_OK;
}

// fused != 0 (real ranks): the traversal block that finishes last also runs the step barrier -- one launch per step
static int sr_traverse(tmd_slabrun *sr, int last, cudaStream_t st, int fused = 0,
                       const cudaGraphConditionalHandle *cond = nullptr) {
    NlArgs a = sr->a;
    const int g = (int)(sr->step & 1), p = sr->p;
    static const int fuse_barrier = nl_env_int("TMD_SR_FUSED_BARRIER", 1);
    const bool in_kernel = fused && fuse_barrier && !sr->langevin && !sr->bonded;
    if (in_kernel) {
        a.bar = sr->step_bar;
        a.bar_gate_out = g ^ 1;
        a.bar_has_cond = cond != nullptr;
        if (cond) a.bar_cond = *cond;
    }
    a.xs = sr->rec_self[p];
    a.xs_next = sr->rec_self[p ^ 1];
    a.peer_lo_next = sr->rec_below[p ^ 1];
    a.peer_hi_next = sr->rec_above[p ^ 1];
    a.flag_clear = sr->gates + g;
    a.flag_next = sr->local_flag;
    a.epi_last = last ? 1 : 0;
    if (sr->langevin) {
        // forces by slot (no integration in the traversal), then ONE pass over the own slots: closing update of this
        // step + (unless this is the last step) opening update of the next, records, halo copies, displacement test
        a.bar = nullptr;
        a.force = sr->s.fs;
        ::tmd::count_launch();
        if (a.ntypes == 1 || a.uniform_table) nl_force_kernel<3, true, 4, true, true, 5, 3><<<sr->fblocks, kNlThreads, 0, st>>>(a);
        else nl_force_kernel<3, false, 4, true, true, 5, 3><<<sr->fblocks, kNlThreads, 0, st>>>(a);
        const int ob = (int)((sr->s.cap_own + 255) / 256);
        ::tmd::count_launch(), sr_langevin_kernel<<<ob, 256, 0, st>>>(
            sr->s, sr->rec_self[p ^ 1], sr->rec_below[p ^ 1], sr->rec_above[p ^ 1], sr->lconsts, sr->noise_key, sr->noise_offset,
            sr->noise_steps, sr->noise_stride, last ? 1 : 3, sr->a.half_skin2, sr->local_flag);
        if (fused) sr_launch_barrier(sr, 3, 3, nullptr, st, cond, last ? 0 : 1);  // (cannot ride in the traversal here)
        TMD_CUDA(cudaGetLastError());
        return TMD_OK;
    }
    if (sr->bonded) {  // the bond's forces at the positions this traversal evaluates: table of parity p -> fx
        a.fx = sr->fx;
        ::tmd::count_launch(), sr_bond_force_kernel<<<1, kSrMaxBonded, 0, st>>>(sr->bond, sr->active_self[p], sr->fx,
                                                                               TMD_WANT_F | (last ? (TMD_WANT_V | TMD_WANT_W) : 0u));
    }
    ::tmd::count_launch();
    if (a.ntypes == 1 || a.uniform_table) nl_force_kernel<3, true, 4, true, true, 5, 2><<<sr->fblocks, kNlThreads, 0, st>>>(a);
    else nl_force_kernel<3, false, 4, true, true, 5, 2><<<sr->fblocks, kNlThreads, 0, st>>>(a);
    if (sr->bonded) {
        if (last)  // System sums the potentials in order: Lennard-Jones wrote this rank's (V, W) share, the bond adds its own
            TMD_CHECK(launch_reduce_vw(sr->bond.partials, 1, 0.5, TMD_WANT_V | TMD_WANT_W | TMD_ACCUMULATE, sr->vw, st));
        else  // the new positions of the bonded atoms this rank owns go to every rank's table of the next parity
            ::tmd::count_launch(), sr_bond_push_kernel<<<1, kSrMaxBonded, 0, st>>>(sr->bond, sr->s.xu, sr->active_all[p ^ 1], sr->world);
    }
    if (fused && !in_kernel) sr_launch_barrier(sr, 3, 3, nullptr, st, cond);
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

/* One piece of a step (tests drive several emulated ranks on one GPU phase by phase: all ranks' signal halves are
 * enqueued before any wait half, so nothing ever spins).  TMD_SR_ADVANCE moves the host-side parities on. */
int tmd_slabrun_phase(tmd_slabrun *sr, int32_t phase, int32_t last, tmd_stream_t stream) {
    TMD_REQUIRE(sr && sr->loaded, "load the system first");
    cudaStream_t st = as_stream(stream);
    const int g = (int)(sr->step & 1);
    const int *gate = sr->gates + g;
    SrArgs s = sr->s;
    s.gate = gate;
    const int ob = (int)((s.cap_own + 255) / 256);
    switch (phase) {
        case TMD_SR_MIGRATE: ::tmd::count_launch(), sr_migrate_kernel<<<ob, 256, 0, st>>>(s); break;
        case TMD_SR_B0_SIGNAL: sr_launch_barrier(sr, 0, 1, gate, st); break;
        case TMD_SR_B0_WAIT: sr_launch_barrier(sr, 0, 2, gate, st); break;
        case TMD_SR_COLLECT: ::tmd::count_launch(), sr_collect_kernel<<<ob, 256, 0, st>>>(s); break;
        case TMD_SR_B1_SIGNAL: sr_launch_barrier(sr, 1, 1, gate, st); break;
        case TMD_SR_B1_WAIT: sr_launch_barrier(sr, 1, 2, gate, st); break;
        case TMD_SR_BUILD: TMD_CHECK(sr_build(sr, gate, 0, st)); break;
        case TMD_SR_B2_SIGNAL: sr_launch_barrier(sr, 2, 1, last ? nullptr : gate, st); break;  // last = 1: ungated (start)
        case TMD_SR_B2_WAIT: sr_launch_barrier(sr, 2, 2, last ? nullptr : gate, st); break;
        case TMD_SR_TRAVERSE: TMD_CHECK(sr_traverse(sr, last, st)); break;
        // (last != 0: the step that ran was a closing one or none at all: no noise was drawn)
        case TMD_SR_B3_SIGNAL: sr_launch_barrier(sr, 3, 1, nullptr, st); break;
        case TMD_SR_B3_WAIT: sr_launch_barrier(sr, 3, 2, nullptr, st, nullptr, sr->langevin && !last ? 1 : 0); break;
        case TMD_SR_PRE: {
            // before step 0: drift to the positions of the first force evaluation.  A run always starts in record
            // buffer 0 and on gate word 0 (so a CUDA graph of two steps recorded after one start fits after any
            // other): the flag of this pass gates the rebuild of step 0, hence the barrier that follows must
            // write gates[0] -- it runs with step = -1.
            sr->p = 0;
            if (sr->langevin)
                ::tmd::count_launch(), sr_langevin_kernel<<<ob, 256, 0, st>>>(
                    s, sr->rec_self[0], sr->rec_below[0], sr->rec_above[0], sr->lconsts, sr->noise_key, sr->noise_offset,
                    sr->noise_steps, sr->noise_stride, 2, sr->a.half_skin2, sr->local_flag);
            else
                ::tmd::count_launch(), sr_pre_kernel<<<ob, 256, 0, st>>>(s, sr->rec_self[0], sr->rec_below[0], sr->rec_above[0],
                                                                       sr->dt, sr->dt * 0.5, sr->a.half_skin2, sr->local_flag);
            if (sr->bonded)
                ::tmd::count_launch(), sr_bond_push_kernel<<<1, kSrMaxBonded, 0, st>>>(sr->bond, sr->s.xu, sr->active_all[0], sr->world);
            sr->step = -1;
            break;
        }
        case TMD_SR_ADVANCE:
            if (sr->step >= 0 && !last) sr->p ^= 1;
            sr->step += 1;
            break;
        default: set_error("tmd_slabrun_phase: unknown phase %d", phase); return TMD_ERR_INVALID;
    }
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

/* LangevinInertia (integrators.py:166-297) instead of VelocityVerlet: constants from tmd_langevin_inertia_constants, noise =
 * the counter-based stream (key, offset = position of the first step's first normal); every step consumes 2 n_global 3
 * normals whatever the number of ranks.  Call before tmd_slabrun_start. */
int tmd_slabrun_set_langevin(tmd_slabrun *sr, const tmd_langevin_consts *consts, uint64_t key, uint64_t offset) {
    TMD_REQUIRE(sr && consts, "NULL argument");
    TMD_REQUIRE(sr->exec == nullptr, "set the integrator before the first run");
    sr->langevin = 1;
    sr->lconsts = *consts;
    sr->noise_key = key;
    sr->noise_offset = offset;
    TMD_CUDA(cudaMemset(sr->noise_steps, 0, sizeof(unsigned long long)));
    return TMD_OK;
}

/* A DoubleWellPair (potentials/well.py:230-286) as the system's second potential (VelocityVerlet runs): gid0 / gid1 = HOST
 * arrays of the global indices of the atoms of types[0] / types[1], ascending (tmd_dwpair's idx0 / idx1); at most 128 atoms
 * in all.  Call before tmd_slabrun_load. */
int tmd_slabrun_set_dwpair(tmd_slabrun *sr, const int32_t *gid0, int32_t n0, const int32_t *gid1, int32_t n1,
                           int32_t same_type, double rwidth, double width2, double height, const tmd_box *box) {
    TMD_REQUIRE(sr && box && box->dim == 3, "NULL argument or not a three-dimensional box");
    TMD_REQUIRE(!sr->loaded && !sr->langevin, "set the bond before tmd_slabrun_load; VelocityVerlet runs only");
    const int nown = same_type ? n0 : n0 + n1;
    TMD_REQUIRE(n0 >= 0 && n1 >= 0 && nown >= 1 && nown <= kSrMaxBonded, "between 1 and 128 atoms of the bonded types");
    TMD_REQUIRE(gid0 || n0 == 0, "gid0 is NULL");
    TMD_REQUIRE(same_type || gid1 || n1 == 0, "gid1 is NULL");
    SrBond &b = sr->bond;
    b.n0 = n0;
    b.n1 = same_type ? 0 : n1;
    b.same_type = same_type ? 1 : 0;
    b.nown = nown;
    b.rwidth = rwidth;
    b.width2 = width2;
    b.height = height;
    for (int k = 0; k < 3; ++k) {
        const bool on = box->periodic[k];
        b.per[k] = on ? 1 : 0;
        b.len[k] = on ? box->length[k] : 0.0;
        b.ilen[k] = on ? box->ilength[k] : 0.0;
        b.half[k] = on ? 0.5 * box->length[k] : 0.0;
    }
    int host[kSrMaxBonded];
    for (int k = 0; k < n0; ++k) host[k] = gid0[k];
    for (int k = 0; k < b.n1; ++k) host[n0 + k] = gid1[k];
    TMD_CUDA(cudaMemcpy(const_cast<int *>(b.gid), host, (size_t)nown * sizeof(int), cudaMemcpyHostToDevice));
    TMD_CUDA(cudaMemset(b.my_slot, 0xFF, kSrMaxBonded * sizeof(int)));
    sr->bonded = 1;
    return TMD_OK;
}

/* after tmd_slabrun_load on every rank: halo addresses, opening half kick + drift, first gate */
int tmd_slabrun_start(tmd_slabrun *sr, tmd_stream_t stream) {
    TMD_CHECK(tmd_slabrun_phase(sr, TMD_SR_B2_SIGNAL, 1, stream));
    TMD_CHECK(tmd_slabrun_phase(sr, TMD_SR_B2_WAIT, 1, stream));
    TMD_CHECK(tmd_slabrun_phase(sr, TMD_SR_PRE, 0, stream));
    TMD_CHECK(tmd_slabrun_phase(sr, TMD_SR_B3_SIGNAL, 0, stream));
    TMD_CHECK(tmd_slabrun_phase(sr, TMD_SR_B3_WAIT, 0, stream));
    return tmd_slabrun_phase(sr, TMD_SR_ADVANCE, 0, stream);
}

/* one whole step on a real rank (its peers run on other GPUs); last != 0: closing half kick only */
int tmd_slabrun_step(tmd_slabrun *sr, int32_t last, tmd_stream_t stream) {
    TMD_REQUIRE(sr && sr->loaded, "load the system first");
    cudaStream_t st = as_stream(stream);
    const int *gate = sr->gates + (int)(sr->step & 1);
    TMD_CHECK(tmd_slabrun_phase(sr, TMD_SR_MIGRATE, 0, stream));
    sr_launch_barrier(sr, 0, 3, gate, st);
    TMD_CHECK(tmd_slabrun_phase(sr, TMD_SR_COLLECT, 0, stream));
    sr_launch_barrier(sr, 1, 3, gate, st);
    TMD_CHECK(tmd_slabrun_phase(sr, TMD_SR_BUILD, 0, stream));
    sr_launch_barrier(sr, 2, 3, gate, st);
    TMD_CHECK(sr_traverse(sr, last, st, 1));
    TMD_CUDA(cudaGetLastError());
    return tmd_slabrun_phase(sr, TMD_SR_ADVANCE, last, stream);
}

// the launches of one rebuild, ungated (inside the IF node of a step graph the decision has been taken)
static int sr_rebuild_ungated(tmd_slabrun *sr, cudaStream_t st) {
    SrArgs s = sr->s;
    s.gate = nullptr;
    const int ob = (int)((s.cap_own + 255) / 256);
    ::tmd::count_launch(), sr_migrate_kernel<<<ob, 256, 0, st>>>(s);
    sr_launch_barrier(sr, 0, 3, nullptr, st);
    ::tmd::count_launch(), sr_collect_kernel<<<ob, 256, 0, st>>>(s);
    sr_launch_barrier(sr, 1, 3, nullptr, st);
    TMD_CHECK(sr_build(sr, nullptr, 0, st));
    sr_launch_barrier(sr, 2, 3, nullptr, st);
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

// Two steps (gate words 0 and 1, record buffers 0 and 1: where every run starts) as one CUDA graph.  The nine
// launches of a rebuild sit in a conditional IF node per step; its condition is set on the device, by the step
// barrier of the step before (or, for the first step of a launch, by a one-thread kernel that reads the gate word
// the previous launch left).  A step without a rebuild is then two kernels: traversal and barrier.
static int sr_build_graph(tmd_slabrun *sr) {
    for (int k = 0; k < 2; ++k)
        if (!sr->capture_stream[k]) TMD_CUDA(cudaStreamCreateWithFlags(&sr->capture_stream[k], cudaStreamNonBlocking));
    cudaStream_t cs = sr->capture_stream[0], bs = sr->capture_stream[1];
    const int64_t step0 = sr->step;
    const int p0 = sr->p;
    TMD_REQUIRE(sr->p == 0 && (sr->step & 1) == 0, "a step graph is recorded right after tmd_slabrun_start");
    TMD_CUDA(cudaGraphCreate(&sr->graph, 0));
    cudaGraphConditionalHandle cond[2];
    for (int k = 0; k < 2; ++k) TMD_CUDA(cudaGraphConditionalHandleCreate(&cond[k], sr->graph, 0, cudaGraphCondAssignDefault));
    const long long launches0 = launch_count_value();
    int status = TMD_OK;
    TMD_CUDA(cudaStreamBeginCaptureToGraph(cs, sr->graph, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    sr_cond_kernel<<<1, 1, 0, cs>>>(cond[0], sr->gates + 0);
    for (int k = 0; k < 2 && status == TMD_OK; ++k) {
        // ---- IF (rebuild) ----
        cudaStreamCaptureStatus cstat;
        cudaGraph_t cg;
        const cudaGraphNode_t *deps;
        size_t ndeps;
        if (cudaStreamGetCaptureInfo_v2(cs, &cstat, nullptr, &cg, &deps, &ndeps) != cudaSuccess) { status = TMD_ERR_CUDA; break; }
        cudaGraphNodeParams np = {};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = cond[k];
        np.conditional.type = cudaGraphCondTypeIf;
        np.conditional.size = 1;
        cudaGraphNode_t node;
        if (cudaGraphAddNode(&node, cg, deps, ndeps, &np) != cudaSuccess) { status = TMD_ERR_CUDA; break; }
        cudaGraph_t body = np.conditional.phGraph_out[0];
        if (cudaStreamUpdateCaptureDependencies(cs, &node, 1, cudaStreamSetCaptureDependencies) != cudaSuccess) { status = TMD_ERR_CUDA; break; }
        if (cudaStreamBeginCaptureToGraph(bs, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { status = TMD_ERR_CUDA; break; }
        status = sr_rebuild_ungated(sr, bs);
        if (cudaStreamEndCapture(bs, nullptr) != cudaSuccess) status = TMD_ERR_CUDA;
        if (status != TMD_OK) break;
        // ---- traversal + step barrier (which decides the next IF) ----
        status = sr_traverse(sr, 0, cs, 1, k == 0 ? &cond[1] : nullptr);
        sr->p ^= 1;
        sr->step += 1;
    }
    cudaGraph_t done = nullptr;
    if (cudaStreamEndCapture(cs, &done) != cudaSuccess) status = TMD_ERR_CUDA;
    sr->step = step0;  // recording is not stepping
    sr->p = p0;
    add_launches_value(launches0 - launch_count_value());
    if (status == TMD_OK && cudaGraphInstantiate(&sr->exec, sr->graph, 0) != cudaSuccess) status = TMD_ERR_CUDA;
    if (status != TMD_OK) {
        cudaGetLastError();
        if (sr->graph) cudaGraphDestroy(sr->graph);
        sr->graph = nullptr;
        sr->exec = nullptr;
        sr->graph_failed = 1;
        set_error("tmd_slabrun: recording the step graph failed; stepping eagerly");
    }
    return status;
}

/* `steps` VelocityVerlet steps (integrators.py:84-94) on a real rank: start, steps - 1 full steps (pairs of them
 * replayed from the step graph when use_graph != 0), one closing step.  Positions and velocities are at the same
 * time afterwards.  graph_launches (optional) = kernels launched through graph replays (the library's launch
 * counter only sees direct launches). */
int tmd_slabrun_run(tmd_slabrun *sr, int32_t steps, int32_t use_graph, int64_t *graph_launches, tmd_stream_t stream) {
    TMD_REQUIRE(sr && sr->loaded, "load the system first");
    TMD_REQUIRE(steps >= 1, "steps must be >= 1");
    cudaStream_t st = as_stream(stream);
    if (graph_launches) *graph_launches = 0;
    TMD_CHECK(tmd_slabrun_start(sr, stream));
    int todo = steps - 1;
    if (use_graph && todo >= 2 && !sr->exec && !sr->graph_failed) sr_build_graph(sr);  // failure: eager steps below
    if (use_graph && sr->exec) {
        while (todo >= 2) {
            TMD_CUDA(cudaGraphLaunch(sr->exec, st));
            sr->step += 2;  // (the buffer parity is back where it was)
            todo -= 2;
            if (graph_launches) *graph_launches += 3;  // a pair without rebuild: condition kernel + 2 traversals
        }
    }
    for (; todo > 0; --todo) TMD_CHECK(tmd_slabrun_step(sr, 0, stream));
    return tmd_slabrun_step(sr, 1, stream);
}

/* own atoms -> their entries of the atom-ordered arrays (n_global x 3; other entries untouched); vw[10] = this
 * rank's share of (V, W) from the last traversal */
int tmd_slabrun_export(tmd_slabrun *sr, double *pos, double *vel, double *force, double *vw, tmd_stream_t stream) {
    TMD_REQUIRE(sr && sr->loaded, "load the system first");
    TMD_REQUIRE((pos && vel && force) || (!pos && !vel && !force && vw), "give pos, vel and force together, or vw alone");
    cudaStream_t st = as_stream(stream);
    const int ob = (int)((sr->s.cap_own + 255) / 256);
    if (pos) ::tmd::count_launch(), sr_export_kernel<<<ob, 256, 0, st>>>(sr->s, pos, vel, force);
    if (vw) TMD_CUDA(cudaMemcpyAsync(vw, sr->vw, 10 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

/* synchronises: error bits (0 = fine), own atoms, list builds so far, layers per rank and along z */
int tmd_slabrun_status(tmd_slabrun *sr, int32_t *err, int64_t *n_own, int64_t *builds, int32_t *layers) {
    TMD_REQUIRE(sr != nullptr, "sr is NULL");
    int e = 0, flags[4] = {0, 0, 0, 0};
    SrCounters c;
    TMD_CUDA(cudaMemcpy(&e, sr->s.self.err, sizeof(int), cudaMemcpyDeviceToHost));
    TMD_CUDA(cudaMemcpy(flags, sr->flags, sizeof(flags), cudaMemcpyDeviceToHost));
    TMD_CUDA(cudaMemcpy(&c, sr->s.cnt, sizeof(c), cudaMemcpyDeviceToHost));
    if (err) *err = e;
    if (n_own) *n_own = c.n_own;
    if (builds) *builds = flags[2];
    if (layers) { layers[0] = sr->s.z1 - sr->s.z0; layers[1] = sr->s.grid.nc[2]; }
    return TMD_OK;
}

}  // extern "C"
