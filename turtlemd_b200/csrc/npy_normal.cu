// K8: numpy.random.Generator(PCG64).standard_normal on the device, bit for bit (npy_normal.h has the restatement and
// the reference call sites it serves: turtlemd/integrators.py:156-160, :216, :312; turtlemd/system/particles.py:260).
//
// The generator is sequential by definition -- normal k starts at the 64-bit word where normal k-1 stopped, and 1.2 % of
// the normals take more than one word -- but PCG64 can jump ahead in O(log) and "which words start a normal" is a chain
// that re-synchronises at once: if a normal is ASSUMED to start at word w, its length is a function of the words from w on
// alone, and 98.8 % of the words are whole normals.  So:
//   canon    one thread per chunk of 32 words jumps to its chunk and walks it as if a normal started at word 0: the
//            normals of that (canonical) chain into the chunk's slot of a scratch buffer, a 32-bit mask of the words
//            where they start and the exit offset, i.e. how many words of the next chunk the last one eats;
//   super    1024 chunks make a super-chunk (one block, records in shared memory), 128 a segment (one warp): lane e
//            follows its segment from entry offset e = 0..31 -- a chain that enters a chunk at a word the canonical chain
//            also starts at IS the canonical chain from there on: count = popc(mask >> e); otherwise (about one chunk
//            in 3000) that chunk is walked again from e -- and leaves (normals produced, exit offset); the first warp
//            then composes the eight segment rows into the row of the super-chunk;
//   chain    one thread runs the super-chunk rows from entry 0 (the stream position): entry offset and number of normals
//            before every super-chunk;
//   place    one block per super-chunk: the true entry through the segment rows, eight threads through their segments
//            (entry offset and output index of every chunk), then every warp moves whole chunks from the scratch buffer
//            to their place in the output -- the canonical normals from the entry word on; a chunk entered off the
//            canonical chain is generated again.  The thread that places normal `count` - 1 leaves the generator state
//            behind it.
// A jump is one 128-bit multiply-add per set bit (a table of the 64 maps "2^k steps ahead" is written with the seed).
// Nothing is read back: the state lives on the device (two slots, flipped by the host per call) until tmd_npstream_state
// asks for it.  An entry offset of 32 or more (one normal eating a whole chunk: probability < 1e-45) or a window that
// turns out too short (the window is count * 1.025 + 4096 words against a measured mean of 1.02204 per normal, > 100
// sigma) raise a sticky status instead of producing a wrong stream.
#if defined(TMD_NPY_EMULATE)
#include "cuda_emul.h"  // tests/emul: the kernels below run as host threads, one block at a time (CPU test of the passes)
#else
#include "common.cuh"
#define TMD_LAUNCH(kernel, grid, block, stream) kernel<<<(grid), (block), 0, (stream)>>>
#endif
#include "npy_normal.h"

namespace tmd {
namespace npy {

constexpr int kChunk = 32;        // words per chunk (= bits of the start mask)
constexpr int kSeg = 128;         // chunks per segment (one warp of the super pass)
constexpr int kSegs = 8;          // segments per super-chunk
constexpr int kSuper = kSeg * kSegs;  // chunks per super-chunk
constexpr int kChainTile = 256;   // super-chunk rows staged per tile of the chain pass
constexpr int kPlaceThreads = 1024;
constexpr unsigned kOffChain = 0x80u;  // entry byte: the chunk is entered off the canonical chain

__device__ const uint64_t g_ki[256] = {TMD_NPY_KI_VALUES};
__device__ const double g_wi[256] = {TMD_NPY_WI_VALUES};
__device__ const double g_fi[256] = {TMD_NPY_FI_VALUES};
static const uint64_t h_ki[256] = {TMD_NPY_KI_VALUES};
static const double h_wi[256] = {TMD_NPY_WI_VALUES};
static const double h_fi[256] = {TMD_NPY_FI_VALUES};

// device words of a stream: state of slot 0 (hi, lo), state of slot 1 (hi, lo), increment (hi, lo), words consumed since
// the seed, status, chunks walked again off the canonical chain by any pass and chunks PLACED off it (coverage counters
// for the tests); then the jump table
// of the generator: (multiplier, increment) of the map "2^k steps ahead", k = 0..63, four words each (written by the seed)
enum { W_STATE0 = 0, W_STATE1 = 2, W_INC = 4, W_CONSUMED = 6, W_STATUS = 7, W_REWALKS = 8, W_OFFCHAIN = 9, W_JUMP = 16,
       W_COUNT = W_JUMP + 4 * 64 };
enum { ST_OK = 0, ST_WINDOW = 1, ST_ENTRY = 2 };

struct Stream {
    int device = 0;
    int slot = 0;  // which state slot is current
    bool seeded = false;
    bool fused = false;
    uint64_t *words = nullptr;
    // scratch, grown on demand
    int64_t cap_chunks = 0;
    uint64_t *recs = nullptr;         // canonical record per chunk: start mask | exit << 32
    double *canon = nullptr;          // canonical normals, 32 slots per chunk
    uint32_t *segtab = nullptr;       // per segment and entry offset: normals | exit offset << 16
    uint32_t *table = nullptr;        // the same per super-chunk
    uint32_t *super_in = nullptr;     // per super-chunk: true entry offset ...
    int64_t *super_prefix = nullptr;  // ... and normals before it
};

struct SmemTables {
    uint64_t ki[256];
    double wi[256], fi[256];
};

// the state `delta` words ahead: one multiply-add of 128 bits per set bit of delta (the maps commute); `jump` is the
// table of the stream, which the kernels keep in shared memory (a jump is a chain of dependent table reads)
__device__ __forceinline__ u128 jump_ahead(const uint64_t *jump, u128 state, uint64_t delta) {
    for (int k = 0; delta != 0; ++k, delta >>= 1) {
        if (delta & 1ull)
            state = state * make_u128(jump[4 * k], jump[4 * k + 1]) + make_u128(jump[4 * k + 2], jump[4 * k + 3]);
    }
    return state;
}

struct SmemJump {
    uint64_t map[4 * 64];
};

__device__ __forceinline__ void load_jump(SmemJump &j, const uint64_t *words) {  // (the caller synchronises)
    for (int k = threadIdx.x; k < 4 * 64; k += blockDim.x) j.map[k] = words[W_JUMP + k];
}

struct SmemBlock {
    SmemTables tab;
    SmemJump jump;
    uint64_t base[2];  // generator state at the first word of the block's first chunk
};

// the ziggurat tables and the jump table into shared memory; thread 0 then jumps to the block's first word, so that
// every thread is a short jump (its chunk within the block) away from its own
__device__ __forceinline__ void load_block(SmemBlock &b, const uint64_t *words, int slot, uint64_t block_word) {
    load_jump(b.jump, words);
    for (int k = threadIdx.x; k < 256; k += blockDim.x) {
        b.tab.ki[k] = g_ki[k];
        b.tab.wi[k] = g_wi[k];
        b.tab.fi[k] = g_fi[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const u128 s = jump_ahead(b.jump.map, make_u128(words[2 * slot], words[2 * slot + 1]), block_word);
        b.base[0] = (uint64_t)(s >> 64);
        b.base[1] = (uint64_t)s;
    }
    __syncthreads();
}

__device__ __forceinline__ Pcg64 generator_in_block(const SmemBlock &b, const uint64_t *words, uint64_t local_word) {
    Pcg64 g;
    g.inc = make_u128(words[W_INC], words[W_INC + 1]);
    g.state = jump_ahead(b.jump.map, make_u128(b.base[0], b.base[1]), local_word);
    return g;
}

__device__ __forceinline__ Pcg64 generator_at(const uint64_t *words, const uint64_t *jump, int slot,
                                              uint64_t word_index) {
    Pcg64 g;
    g.inc = make_u128(words[W_INC], words[W_INC + 1]);
    g.state = jump_ahead(jump, make_u128(words[2 * slot], words[2 * slot + 1]), word_index);
    return g;
}

// four normals to out[0..3], out 32-byte aligned: one 256-bit store (a thread fills its chunk's slot front to back, the
// lanes of a warp are 256 bytes apart: every store instruction is one transaction per lane whatever its width)
__device__ __forceinline__ void store4(double *out, double a, double b, double c, double d) {
#if defined(TMD_NPY_EMULATE)
    out[0] = a;
    out[1] = b;
    out[2] = c;
    out[3] = d;
#else
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(out), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#endif
}

// Walk one chunk from word `entry`: the start mask of the chain, how many normals start in the chunk and how far the
// last one reaches into the next chunk.  EMIT stores the normals (out[0 .. count-1]); `stop_at` >= 0 makes the walk
// report the number of words consumed up to and including its normal number `stop_at` (0-based) in `words_at_stop`.
// The inner loop takes the words that are accepted on the spot -- all lanes of a warp in lock step -- and is left at a
// wedge or tail word; the lanes that have one finish it together behind the loop (zig_slow is 5-10x a fast word, and
// with one chunk per lane some lane meets one in a third of the iterations).
enum { EMIT_NONE = 0, EMIT_EACH = 1, EMIT_BY_FOUR = 2 };  // EMIT_BY_FOUR: out is global and 32-byte aligned (store4)

template <int EMIT>
__device__ __forceinline__ void emit_normal(double *out, uint32_t count, double x, double (&q)[4]) {
    if (EMIT == EMIT_EACH) out[count] = x;
    if (EMIT == EMIT_BY_FOUR) {  // (selects, not an indexed array: the group stays in registers)
        const uint32_t j = count & 3u;
        q[0] = j == 0u ? x : q[0];
        q[1] = j == 1u ? x : q[1];
        q[2] = j == 2u ? x : q[2];
        if (j == 3u) store4(out + (count - 3u), q[0], q[1], q[2], x);
    }
}

template <int EMIT>
__device__ __forceinline__ void walk_chunk(Pcg64 g, const ZigTables &t, uint32_t entry, uint32_t &mask, uint32_t &count,
                                           uint32_t &exit_offset, double *out, int64_t stop_at, uint32_t &words_at_stop) {
    uint32_t pos = entry;
    double q[4] = {0.0, 0.0, 0.0, 0.0};
    mask = 0;
    count = 0;
    while (pos < (uint32_t)kChunk) {
        uint64_t r = 0;
        double x = 0.0;
        bool pending = false;
        while (pos < (uint32_t)kChunk) {
            r = pcg_next(g);
            mask |= 1u << pos;
            ++pos;
            if (!zig_fast(r, t, x)) {
                pending = true;
                break;
            }
            if (EMIT != EMIT_NONE) {
                emit_normal<EMIT>(out, count, x, q);
                if ((int64_t)count == stop_at) words_at_stop = pos;
            }
            ++count;
        }
        if (pending) {
            uint32_t used = 0;
            x = zig_slow(g, t, r, x, used);
            pos += used;
            if (EMIT != EMIT_NONE) {
                emit_normal<EMIT>(out, count, x, q);
                if ((int64_t)count == stop_at) words_at_stop = pos;
            }
            ++count;
        }
    }
    if (EMIT == EMIT_BY_FOUR && (count & 3u) != 0u)  // the last, partial group (the slot holds 32 normals: room enough)
        store4(out + (count & ~3u), q[0], q[1], q[2], q[3]);
    exit_offset = pos - (uint32_t)kChunk;
}

__global__ void __launch_bounds__(128) np_canon_kernel(const uint64_t *words, int slot, bool fused,
                                                       uint64_t *__restrict__ recs, double *__restrict__ canon,
                                                       int64_t n_chunks) {
    __shared__ SmemBlock sm;
    load_block(sm, words, slot, (uint64_t)blockIdx.x * blockDim.x * kChunk);
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const ZigTables t = {sm.tab.ki, sm.tab.wi, sm.tab.fi, fused};
    uint32_t mask, count, exit_offset, unused = 0;
    walk_chunk<EMIT_BY_FOUR>(generator_in_block(sm, words, (uint64_t)threadIdx.x * kChunk), t, 0u, mask, count,
                             exit_offset, canon + c * kChunk, -1, unused);
    recs[c] = (uint64_t)mask | ((uint64_t)exit_offset << 32);
}

// a chunk entered at a word the canonical chain does not start at is walked again from there (about one chunk in 3000);
// returns normals produced | exit offset << 16
__device__ __noinline__ uint32_t rewalk_chunk(uint64_t *words, const uint64_t *jump, int slot, bool fused,
                                              int64_t chunk, uint32_t e) {
    const ZigTables t = {g_ki, g_wi, g_fi, fused};
    uint32_t mask, produced, exit_offset, unused = 0;
    atomicAdd(reinterpret_cast<unsigned long long *>(words + W_REWALKS), 1ull);
    walk_chunk<EMIT_NONE>(generator_at(words, jump, slot, (uint64_t)chunk * kChunk + e), t, e, mask, produced,
                          exit_offset, nullptr, -1, unused);
    if (exit_offset >= (uint32_t)kChunk) {  // a normal longer than a chunk: flagged, the stream is not used
        words[W_STATUS] = ST_ENTRY;
        exit_offset = kChunk - 1;
    }
    return produced | (exit_offset << 16);
}

// one chunk of a chain that enters it at word e < 32: normals produced | exit offset << 16 (| kOffChain << 24 when the
// chunk had to be walked again)
__device__ __forceinline__ uint32_t follow_chunk(uint64_t *words, const uint64_t *jump, int slot, bool fused,
                                                 uint64_t rec, int64_t chunk, uint32_t e) {
    const uint32_t mask = (uint32_t)rec;
    if ((mask >> e) & 1u) {  // on the canonical chain from here on (always so for e == 0)
        uint32_t exit_offset = (uint32_t)(rec >> 32);
        if (exit_offset >= (uint32_t)kChunk) {
            words[W_STATUS] = ST_ENTRY;
            exit_offset = kChunk - 1;
        }
        return (uint32_t)__popc(mask >> e) | (exit_offset << 16);
    }
    return rewalk_chunk(words, jump, slot, fused, chunk, e) | (kOffChain << 24);
}

// one block per super-chunk: warp w, lane e follows segment w from entry e; then lane e of warp 0 composes the segments
__global__ void __launch_bounds__(256) np_super_kernel(uint64_t *words, int slot, bool fused,
                                                       const uint64_t *__restrict__ recs, uint32_t *__restrict__ segtab,
                                                       uint32_t *__restrict__ table) {
    __shared__ uint64_t rec_s[kSuper];
    __shared__ uint32_t seg_s[kSegs * kChunk];
    __shared__ SmemJump jump_s;
    const int64_t c0 = (int64_t)blockIdx.x * kSuper;
    for (int k = threadIdx.x; k < kSuper; k += blockDim.x) rec_s[k] = recs[c0 + k];
    load_jump(jump_s, words);
    __syncthreads();
    {
        const int w = threadIdx.x / kChunk;
        uint32_t e = threadIdx.x % kChunk, produced = 0;
        for (int k = w * kSeg; k < (w + 1) * kSeg; ++k) {
            const uint32_t step = follow_chunk(words, jump_s.map, slot, fused, rec_s[k], c0 + k, e);
            produced += step & 0xffffu;
            e = (step >> 16) & 0xffu;
        }
        const uint32_t row = produced | (e << 16);  // produced <= 128 * 32
        seg_s[threadIdx.x] = row;
        segtab[(int64_t)blockIdx.x * (kSegs * kChunk) + threadIdx.x] = row;
    }
    __syncthreads();
    if (threadIdx.x < (unsigned)kChunk) {
        uint32_t e = threadIdx.x, produced = 0;
        for (int w = 0; w < kSegs; ++w) {
            const uint32_t row = seg_s[w * kChunk + e];
            produced += row & 0xffffu;
            e = row >> 16;
        }
        table[(int64_t)blockIdx.x * kChunk + threadIdx.x] = produced | (e << 16);  // produced <= 32768
    }
}

// one block: the rows of the super-chunks staged a tile at a time, thread 0 runs the chain from entry 0
__global__ void __launch_bounds__(1024) np_chain_kernel(uint64_t *words, const uint32_t *__restrict__ table,
                                                        uint32_t *__restrict__ super_in,
                                                        int64_t *__restrict__ super_prefix, int64_t n_supers,
                                                        int64_t count) {
    __shared__ uint32_t rows[kChainTile * kChunk];
    uint32_t e = 0;
    int64_t p = 0;
    for (int64_t base = 0; base < n_supers; base += kChainTile) {
        const int64_t here = (n_supers - base < kChainTile) ? n_supers - base : kChainTile;
        __syncthreads();
        for (int64_t k = threadIdx.x; k < here * kChunk; k += blockDim.x) rows[k] = table[base * kChunk + k];
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int64_t s = 0; s < here; ++s) {
                super_in[base + s] = e;
                super_prefix[base + s] = p;
                const uint32_t row = rows[s * kChunk + e];
                p += row & 0xffffu;
                e = row >> 16;
            }
        }
    }
    if (threadIdx.x == 0) {
        super_prefix[n_supers] = p;
        if (p < count && words[W_STATUS] == ST_OK) words[W_STATUS] = ST_WINDOW;
    }
}

// the word behind canonical normal number q of a chunk: the start of canonical normal q + 1, or the chunk's end + exit
__device__ __forceinline__ uint32_t word_behind(uint64_t rec, uint32_t q) {
    uint32_t m = (uint32_t)rec;
    for (uint32_t k = 0; k <= q; ++k) m &= m - 1;  // drop the starts of normals 0 .. q
    return m ? (uint32_t)(__ffs((int)m) - 1) : (uint32_t)kChunk + (uint32_t)(rec >> 32);
}

__global__ void __launch_bounds__(kPlaceThreads) np_place_kernel(uint64_t *words, int slot, bool fused,
                                                                 const uint64_t *__restrict__ recs,
                                                                 const double *__restrict__ canon,
                                                                 const uint32_t *__restrict__ segtab,
                                                                 const uint32_t *__restrict__ super_in,
                                                                 const int64_t *__restrict__ super_prefix,
                                                                 double *__restrict__ out, int64_t count) {
    __shared__ uint64_t rec_s[kSuper];
    __shared__ uint32_t seg_s[kSegs * kChunk];
    __shared__ uint32_t rel_s[kSuper + 1];     // normals of this super-chunk before each chunk (+ the total)
    __shared__ unsigned char entry_s[kSuper];  // true entry offset | kOffChain
    __shared__ uint32_t seg_entry[kSegs], seg_rel[kSegs + 1];
    __shared__ double again_s[kPlaceThreads / 32][kChunk];  // per warp: a chunk generated again
    __shared__ SmemJump jump_s;
    const int64_t p0 = super_prefix[blockIdx.x];
    if (p0 >= count) return;  // the whole super-chunk lies behind the last normal asked for
    const int64_t c0 = (int64_t)blockIdx.x * kSuper;
    for (int k = threadIdx.x; k < kSuper; k += blockDim.x) rec_s[k] = recs[c0 + k];
    if (threadIdx.x < (unsigned)(kSegs * kChunk))
        seg_s[threadIdx.x] = segtab[(int64_t)blockIdx.x * (kSegs * kChunk) + threadIdx.x];
    load_jump(jump_s, words);
    __syncthreads();
    if (threadIdx.x == 0) {  // the true entry through the segment rows
        uint32_t e = super_in[blockIdx.x], produced = 0;
        for (int w = 0; w < kSegs; ++w) {
            seg_entry[w] = e;
            seg_rel[w] = produced;
            const uint32_t row = seg_s[w * kChunk + e];
            produced += row & 0xffffu;
            e = row >> 16;
        }
        seg_rel[kSegs] = produced;
        rel_s[kSuper] = produced;
    }
    __syncthreads();
    if ((threadIdx.x & 31u) == 0 && threadIdx.x / 32 < (unsigned)kSegs) {  // one thread per segment, each in its own warp
        const int w = threadIdx.x / 32;
        uint32_t e = seg_entry[w], produced = seg_rel[w];
        for (int k = w * kSeg; k < (w + 1) * kSeg; ++k) {
            const uint32_t step = follow_chunk(words, jump_s.map, slot, fused, rec_s[k], c0 + k, e);
            entry_s[k] = (unsigned char)(e | (step >> 24));
            rel_s[k] = produced;
            produced += step & 0xffffu;
            e = (step >> 16) & 0xffu;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    // every warp moves whole chunks, lane j normal j of a chunk, four chunks in flight: the canonical normals from the
    // entry word on
    for (int k0 = warp; k0 < kSuper; k0 += 4 * (kPlaceThreads / 32)) {
        double v[4];
        int64_t dst[4];
        bool move[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k0 + u * (kPlaceThreads / 32);
            move[u] = false;
            if (k < kSuper) {
                const int64_t p = p0 + rel_s[k];
                const uint32_t n = rel_s[k + 1] - rel_s[k];
                const uint32_t e = entry_s[k];
                if (!(e & kOffChain) && (uint32_t)lane < n && p + lane < count) {
                    const uint32_t skip = (uint32_t)__popc((uint32_t)rec_s[k] & ((1u << e) - 1u));  // normals before word e
                    v[u] = canon[(c0 + k) * kChunk + skip + lane];
                    dst[u] = p + lane;
                    move[u] = true;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (move[u]) out[dst[u]] = v[u];
    }
    // the rare rest: chunks entered off the canonical chain are generated again, and the chunk that holds normal
    // `count` - 1 leaves the generator state behind it
    for (int k = warp; k < kSuper; k += kPlaceThreads / 32) {
        const int64_t p = p0 + rel_s[k];
        if (p >= count) break;  // (the same for all lanes; later chunks of this warp lie further behind)
        const uint32_t n = rel_s[k + 1] - rel_s[k];
        const bool off_chain = (entry_s[k] & kOffChain) != 0;
        const bool last_here = count - 1 - p < (int64_t)n;  // normal count - 1 is one of this chunk's
        if (!off_chain && !last_here) continue;
        const uint32_t e = entry_s[k] & 31u;
        const uint64_t rec = rec_s[k];
        uint32_t behind = 0;  // words of the chunk consumed up to and including normal count - 1
        if (off_chain) {
            if (lane == 0) {
                const ZigTables t = {g_ki, g_wi, g_fi, fused};
                uint32_t mask, produced, exit_offset;
                atomicAdd(reinterpret_cast<unsigned long long *>(words + W_OFFCHAIN), 1ull);
                walk_chunk<EMIT_EACH>(generator_at(words, jump_s.map, slot, (uint64_t)(c0 + k) * kChunk + e), t, e, mask,
                                      produced, exit_offset, again_s[warp], count - 1 - p, behind);
            }
            __syncwarp();
            if ((uint32_t)lane < n && p + lane < count) out[p + lane] = again_s[warp][lane];
            __syncwarp();
        } else if (lane == 0) {
            const uint32_t skip = (uint32_t)__popc((uint32_t)rec & ((1u << e) - 1u));
            behind = word_behind(rec, skip + (uint32_t)(count - 1 - p));
        }
        if (last_here && lane == 0 && words[W_STATUS] == ST_OK) {
            const uint64_t consumed = (uint64_t)(c0 + k) * kChunk + behind;
            const u128 next = jump_ahead(jump_s.map, make_u128(words[2 * slot], words[2 * slot + 1]), consumed);
            words[2 * (slot ^ 1)] = (uint64_t)(next >> 64);
            words[2 * (slot ^ 1) + 1] = (uint64_t)next;
            words[W_CONSUMED] += consumed;
        }
    }
}

// the seed: state, increment and the jump table (multiplier, increment) of 2^k steps, k = 0..63
__global__ void np_seed_kernel(uint64_t *words, int slot, uint64_t s_hi, uint64_t s_lo, uint64_t i_hi, uint64_t i_lo) {
    words[2 * slot] = s_hi;
    words[2 * slot + 1] = s_lo;
    words[W_INC] = i_hi;
    words[W_INC + 1] = i_lo;
    words[W_CONSUMED] = 0;
    words[W_STATUS] = ST_OK;
    words[W_REWALKS] = 0;
    words[W_OFFCHAIN] = 0;
    u128 mult = pcg_multiplier(), plus = make_u128(i_hi, i_lo);
    for (int k = 0; k < 64; ++k) {
        words[W_JUMP + 4 * k] = (uint64_t)(mult >> 64);
        words[W_JUMP + 4 * k + 1] = (uint64_t)mult;
        words[W_JUMP + 4 * k + 2] = (uint64_t)(plus >> 64);
        words[W_JUMP + 4 * k + 3] = (uint64_t)plus;
        plus = (mult + 1) * plus;
        mult *= mult;
    }
}

static int64_t window_chunks(int64_t count) {
    const int64_t words = count + count / 40 + 4096;
    const int64_t per_super = (int64_t)kChunk * kSuper;
    return ((words + per_super - 1) / per_super) * kSuper;
}

static void release_scratch(Stream *h) {
    cudaFree(h->recs);
    cudaFree(h->canon);
    cudaFree(h->segtab);
    cudaFree(h->table);
    cudaFree(h->super_in);
    cudaFree(h->super_prefix);
    h->recs = nullptr;
    h->canon = nullptr;
    h->segtab = nullptr;
    h->table = nullptr;
    h->super_in = nullptr;
    h->super_prefix = nullptr;
    h->cap_chunks = 0;
}

static int reserve(Stream *h, int64_t n_chunks, cudaStream_t stream) {
    if (n_chunks <= h->cap_chunks) return TMD_OK;
    TMD_CUDA(cudaStreamSynchronize(stream));  // launches of earlier calls may still read the old scratch
    release_scratch(h);
    const int64_t want = n_chunks + n_chunks / 8;
    const int64_t chunks = ((want + kSuper - 1) / kSuper) * kSuper, supers = chunks / kSuper;
    cudaError_t err = cudaMalloc(&h->recs, sizeof(uint64_t) * chunks);
    if (err == cudaSuccess) err = cudaMalloc(&h->canon, sizeof(double) * chunks * kChunk);
    if (err == cudaSuccess) err = cudaMalloc(&h->segtab, sizeof(uint32_t) * supers * kSegs * kChunk);
    if (err == cudaSuccess) err = cudaMalloc(&h->table, sizeof(uint32_t) * supers * kChunk);
    if (err == cudaSuccess) err = cudaMalloc(&h->super_in, sizeof(uint32_t) * supers);
    if (err == cudaSuccess) err = cudaMalloc(&h->super_prefix, sizeof(int64_t) * (supers + 1));
    if (err != cudaSuccess) {
        release_scratch(h);
        return cuda_fail(err, "cudaMalloc (numpy stream scratch)", __FILE__, __LINE__);
    }
    h->cap_chunks = chunks;
    return TMD_OK;
}

// which log1p NumPy's process runs: libm's own answers on twelve arguments where the two builds differ in the last bit
// (found by running both restatements over 4e6 uniform arguments); 1 = the FMA build, 0 = the baseline build, -1 = neither
static int libm_variant() {
    static int cached = -2;
    if (cached != -2) return cached;
    static const uint64_t probes[12] = {
        0xBFEFAB720963D8A9ull, 0xBFE500819F0ED9A2ull, 0xBFDA8EAF9AD6259Aull, 0xBFD636AAA46F6AF6ull,
        0xBFD439B24026C5DAull, 0xBFD357C889E5D6A0ull, 0xBFD25E83648AA84Eull, 0xBFD17C24F95A1932ull,
        0xBFCFC8D3ABE09D28ull, 0xBFCB4FD4CAD991ECull, 0xBFC6B0614FC5DEB0ull, 0xBFA511FAD4170590ull};
    bool plain_ok = true, fused_ok = true;
    for (uint64_t bits : probes) {
        volatile double x = u64_as_double(bits);
        const uint64_t ref = double_as_u64(log1p((double)x));
        if (double_as_u64(log1p_glibc(x, false)) != ref) plain_ok = false;
        if (double_as_u64(log1p_glibc(x, true)) != ref) fused_ok = false;
    }
    cached = fused_ok ? 1 : (plain_ok ? 0 : -1);
    return cached;
}

}  // namespace npy
}  // namespace tmd

using namespace tmd;
using namespace tmd::npy;

extern "C" {

int tmd_npstream_libm_variant(void) { return libm_variant(); }

int tmd_npstream_create(void **handle) {
    TMD_REQUIRE(handle != nullptr, "handle is NULL");
    TMD_CHECK(require_device());
    const int variant = libm_variant();
    if (variant < 0) {
        set_error("tmd_npstream_create: this libm's log1p is neither of the two glibc builds restated in npy_normal.h; "
                  "the device stream could not match NumPy's tail draws");
        return TMD_ERR_INVALID;
    }
    Stream *h = new Stream();
    h->fused = variant == 1;
    TMD_CUDA(cudaGetDevice(&h->device));
    cudaError_t err = cudaMalloc(&h->words, sizeof(uint64_t) * W_COUNT);
    if (err != cudaSuccess) {
        delete h;
        return cuda_fail(err, "cudaMalloc (numpy stream state)", __FILE__, __LINE__);
    }
    *handle = h;
    return TMD_OK;
}

int tmd_npstream_destroy(void *handle) {
    Stream *h = static_cast<Stream *>(handle);
    if (!h) return TMD_OK;
    cudaDeviceSynchronize();
    release_scratch(h);
    cudaFree(h->words);
    delete h;
    return TMD_OK;
}

int tmd_npstream_seed(void *handle, const uint64_t *state_inc, tmd_stream_t stream) {
    Stream *h = static_cast<Stream *>(handle);
    TMD_REQUIRE(h != nullptr && state_inc != nullptr, "handle / state is NULL");
    TMD_REQUIRE((state_inc[3] & 1ull) == 1ull, "the PCG64 increment must be odd");
    count_launch();
    TMD_LAUNCH(np_seed_kernel, 1, 1, as_stream(stream))(h->words, h->slot, state_inc[0], state_inc[1], state_inc[2],
                                                   state_inc[3]);
    TMD_CUDA(cudaGetLastError());
    h->seeded = true;
    return TMD_OK;
}

int tmd_npstream_normals(void *handle, double *out, int64_t count, tmd_stream_t stream) {
    Stream *h = static_cast<Stream *>(handle);
    TMD_REQUIRE(h != nullptr && h->seeded, "stream is NULL or was never seeded");
    TMD_REQUIRE(count >= 0 && (count == 0 || out != nullptr), "count must be >= 0 and out non-NULL");
    cudaStream_t s = as_stream(stream);
    const int64_t kMaxPerCall = (int64_t)1 << 23;  // bounds the scratch (264 B per 32 words) and the serial chain pass
    for (int64_t done = 0; done < count;) {
        const int64_t now = (count - done < kMaxPerCall) ? count - done : kMaxPerCall;
        const int64_t n_chunks = window_chunks(now), n_supers = n_chunks / kSuper;
        TMD_CHECK(reserve(h, n_chunks, s));
        count_launch();
        TMD_LAUNCH(np_canon_kernel, (unsigned)((n_chunks + 127) / 128), 128, s)(h->words, h->slot, h->fused, h->recs,
                                                                                  h->canon, n_chunks);
        count_launch();
        TMD_LAUNCH(np_super_kernel, (unsigned)n_supers, 256, s)(h->words, h->slot, h->fused, h->recs, h->segtab, h->table);
        count_launch();
        TMD_LAUNCH(np_chain_kernel, 1, 1024, s)(h->words, h->table, h->super_in, h->super_prefix, n_supers, now);
        count_launch();
        TMD_LAUNCH(np_place_kernel, (unsigned)n_supers, kPlaceThreads, s)(h->words, h->slot, h->fused, h->recs, h->canon,
                                                                            h->segtab, h->super_in, h->super_prefix,
                                                                            out + done, now);
        TMD_CUDA(cudaGetLastError());
        h->slot ^= 1;
        done += now;
    }
    return TMD_OK;
}

int tmd_npstream_state(void *handle, uint64_t *state, uint64_t *words_consumed, uint64_t *counters,
                       tmd_stream_t stream) {
    Stream *h = static_cast<Stream *>(handle);
    TMD_REQUIRE(h != nullptr && h->seeded, "stream is NULL or was never seeded");
    uint64_t host[W_JUMP];  // everything but the jump table
    TMD_CUDA(cudaMemcpyAsync(host, h->words, sizeof(host), cudaMemcpyDeviceToHost, as_stream(stream)));
    TMD_CUDA(cudaStreamSynchronize(as_stream(stream)));
    if (host[W_STATUS] != ST_OK) {
        set_error("tmd_npstream: %s", host[W_STATUS] == ST_WINDOW
                                           ? "the word window of a call was too short for the normals asked for"
                                           : "a single normal consumed more than a chunk of 32 words");
        return TMD_ERR_INVALID;
    }
    if (state) {
        state[0] = host[2 * h->slot];
        state[1] = host[2 * h->slot + 1];
    }
    if (words_consumed) *words_consumed = host[W_CONSUMED];
    if (counters) {
        counters[0] = host[W_REWALKS];
        counters[1] = host[W_OFFCHAIN];
    }
    return TMD_OK;
}

int tmd_host_npstream_normals(uint64_t *state_inc, double *out, int64_t count, uint64_t *words_consumed) {
    TMD_REQUIRE(state_inc != nullptr && count >= 0 && (count == 0 || out != nullptr), "bad arguments");
    const int variant = libm_variant();
    TMD_REQUIRE(variant >= 0, "this libm's log1p is not one of the glibc builds restated in npy_normal.h");
    Pcg64 g = {make_u128(state_inc[0], state_inc[1]), make_u128(state_inc[2], state_inc[3])};
    const ZigTables t = {h_ki, h_wi, h_fi, variant == 1};
    uint64_t total = 0;
    for (int64_t k = 0; k < count; ++k) {
        uint32_t used = 0;
        out[k] = standard_normal(g, t, used);
        total += used;
    }
    state_inc[0] = (uint64_t)(g.state >> 64);
    state_inc[1] = (uint64_t)g.state;
    if (words_consumed) *words_consumed = total;
    return TMD_OK;
}

#if defined(TMD_NPY_EMULATE)
// test hooks of the host emulation build only (tests/test_npy_normal.py)
void emul_log1p(const double *x, double *y, int64_t n, int32_t fused) {
    for (int64_t k = 0; k < n; ++k) y[k] = log1p_glibc(x[k], fused != 0);
}
void emul_pcg_advance(uint64_t *state_inc, uint64_t delta) {
    const u128 s = pcg_advance(make_u128(state_inc[0], state_inc[1]), make_u128(state_inc[2], state_inc[3]), delta);
    state_inc[0] = (uint64_t)(s >> 64);
    state_inc[1] = (uint64_t)s;
}
#endif

}  // extern "C"
