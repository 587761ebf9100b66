// Read-sketch kernel (hand-written sm_100a CUDA), 256-thread tiles.
//
//   read_sketch_kernel : ILP_index::compute_hashes for every read + the Sp_R union
//                        (/root/reference/src/ILP_index.cpp:447-493, :615-629), fused:
//                        minimizer hashes go straight into an order-preserving open-addressing HBM table.
#define PHI_TILE_THREADS 256
#include "kernels.h"
#include "sketch_common.cuh"

namespace phi {
static_assert(PHI_TILE_THREADS == READ_TILE_THREADS, "tile size");

// FAST: register-resident core (CLEAN tile, 9 <= w <= 65); otherwise the shared-memory core (any w, any byte)
template <bool CLEAN, bool FAST>
__device__ __forceinline__ void read_tile_body(Tile &t, const ReadSketchArgs &A)
{
    const int tid = threadIdx.x;
    uint16_t *runs = t.pre;                                          // safe: both cores sync before writing runs[]
    uint64_t *run_val = t.canon;                                     // FAST only (canon[] is not used there)
    int halo = -1, n_runs; uint64_t halo_val = 0;
    if (FAST) {
        n_runs = fast_runs<true>(t, runs, run_val, &halo, &halo_val);
    } else {
        phase_canon<CLEAN>(t);
        __syncthreads();
        phase_block_minima<CLEAN>(t);
        __syncthreads();
        n_runs = phase_runs<true, CLEAN>(t, runs, &halo);
    }
    if (n_runs == 0) return;

    if (FAST) { if (tid == 32 * 0 + tile_halo_lanes(t.w)) t.hash[0] = halo >= 0 ? hash_packed_kmer(halo_val, t.k) : 0xFFFFFFFFFFFFFFFFull; }
    else if (tid == 0) t.hash[0] = halo >= 0 ? hash_at<CLEAN>(t, halo) : 0xFFFFFFFFFFFFFFFFull;
    int emitted = 0;
    for (int b0 = 0; b0 < n_runs; b0 += NT) {
        const int j = b0 + tid; const bool have = j < n_runs;
        const int cnt = min(NT, n_runs - b0);
        uint32_t ent = have ? runs[j] : 0;
        uint64_t h = 0;
        if (have) h = FAST ? hash_packed_kmer(run_val[j], t.k) : hash_at<CLEAN>(t, ent & 0x7FFF);
        t.hash[tid + 1] = h;
        __syncthreads();
        uint64_t prev = (ent & 0x8000) ? 0xFFFFFFFFFFFFFFFFull : t.hash[tid];
        uint64_t carry = t.hash[cnt];
        bool emit = have && h != prev;
        emitted += __syncthreads_count(emit);                         // everybody has read hash[]: the next batch may overwrite it
        if (tid == 0) t.hash[0] = carry;
        if (emit) table_insert(A.table, A.table_mult, A.table_limit, h, A.ctr);   // after the barrier: nobody waits for the table round trip
    }
    if (tid == 0 && emitted) atomicAdd(&A.ctr[CTR_READ_EMITTED], (unsigned long long)emitted);
}

__global__ void __launch_bounds__(NT, 4)
read_sketch_kernel(ReadSketchArgs A)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const TileLayout &L = A.layout;
    Tile t = carve(smem, L, A.k, A.w);
    const long long tile = (long long)A.tile0 + blockIdx.x;
    t.g0 = tile * L.cap - A.w - L.pad;
    t.seq_len = (long long)A.total_bases;
    set_window_bounds(t);
    const int tid = threadIdx.x;

    // ---- optional: the tile's bytes are contiguous in the read buffer, so ONE bulk copy (cp.async.bulk, completion on an mbarrier)
    // brings them to shared memory while the boundary pass below runs; source, destination and size are multiples of 16 bytes
    // (the buffer has 16 readable bytes in front and 48 behind the reads).
    __shared__ __align__(8) unsigned long long s_mbar;
    unsigned char *raw = smem + L.o_raw;
    long long a0 = 0; bool bulk = A.bulk != 0;
    if (bulk) {
        a0 = t.g0 & ~15ll; if (a0 < -16) a0 = -16;
        long long a1 = (t.g0 + L.NB + 15) & ~15ll;
        const long long lim = ((long long)A.total_bases + 48) & ~15ll;
        if (a1 > lim) a1 = lim;
        const uint32_t bytes = (uint32_t)(a1 - a0);
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_mbar), dst = (uint32_t)__cvta_generic_to_shared(raw);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst), "l"(A.read_bases + a0), "r"(bytes), "r"(bar) : "memory");
        }
    }

    // the first staging load of every thread is issued here, ahead of the boundary pass: its latency and the boundary pass's chain
    // (tile directory -> read offsets) overlap instead of queueing up
    uint64_t v_first = 0;
    const bool have_first = !bulk;
    if (have_first && tid < L.nchunks) {
        const long long g = t.g0 + 8ll * tid;
        if (g + 8 > 0 && g < t.seq_len) v_first = load8_unaligned(A.read_bases + g);
    }

    // ---- read boundaries -> window masks.  A read starting at local base b makes the windows e with b inside their bases
    // (e-w+1, e+k-1], i.e. e in [b-k+1, b+w-2], invalid, and e = b+w-1 the first window of that read.
    const int nwords = (L.M + 31) / 32 + 2;
    for (int i = tid; i < nwords; i += NT) { t.inval[i] = 0; t.firstm[i] = 0; }
    __syncthreads();
    {
        const long long hi = t.g0 + L.NB;
        const int nbits = 32 * nwords;
        uint64_t r0 = A.tile_first_read[tile];
        for (;;) {
            uint64_t r = r0 + tid;
            int past = 1;
            if (r <= A.n_reads) {
                long long off = (long long)A.read_off[r];
                if (off < hi) {
                    past = 0;
                    const int b = (int)(off - t.g0);
                    const int lo = max(b - A.k + 1, 0), hi_b = min(b + A.w - 2, nbits - 1);
                    for (int wi = lo >> 5; wi <= (hi_b >> 5) && lo <= hi_b; ++wi) {
                        uint32_t m = 0xFFFFFFFFu;
                        if (wi == (lo >> 5)) m &= 0xFFFFFFFFu << (lo & 31);
                        if (wi == (hi_b >> 5)) m &= 0xFFFFFFFFu >> (31 - (hi_b & 31));
                        atomicOr(&t.inval[wi], m);
                    }
                    const int f = b + A.w - 1;
                    if (f < nbits) atomicOr(&t.firstm[f >> 5], 1u << (f & 31));
                }
            }
            if (__syncthreads_or(past)) break;
            r0 += NT;
        }
    }
    // ---- stage bases: unaligned 8-byte loads (from the landing zone of the bulk copy, or straight from global memory), mask outside [0, total)
    if (bulk) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
        uint32_t done = 0;
        for (int it = 0; it < (1 << 22) && !done; ++it)              // phase 0 of the barrier completes when all bytes have landed
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar) : "memory");
        if (!__syncthreads_and(done)) bulk = false;                  // (never seen; the direct loads below are always right)
    }
    uint32_t dirty_any = 0;
    for (int c = tid; c < L.nchunks; c += NT) {
        long long g = t.g0 + 8ll * c;
        uint64_t v = 0;
        if (g + 8 > 0 && g < t.seq_len) {
            v = bulk ? load8_unaligned(raw + (g - a0)) : (have_first && c == tid) ? v_first : load8_unaligned(A.read_bases + g);   // front padding covers g in [-7, -1]
            if (g < 0) v &= ~0ull << (8 * (int)(-g));
            long long nvalid = t.seq_len - g;                        // bytes [0, nvalid) of the chunk are real
            if (nvalid < 8) v &= (1ull << (8 * nvalid)) - 1;
            v = upcase8(v);
        }
        dirty_any |= stage_chunk(t, c, v);
    }
    // a tile is CLEAN when every staged byte that a valid window can touch is A/C/G/T; padding at either end of
    // the data counts as dirty and sends the (few) boundary tiles through the general path
    if (__syncthreads_or(dirty_any != 0 || A.k > MAX_PACKED_K)) read_tile_body<false, false>(t, A);
    else if (tile_fast_w(A.w)) read_tile_body<true, true>(t, A);
    else read_tile_body<true, false>(t, A);
}

// per tile: first read r with read_off[r] >= g0 = tile*cap - w - pad
__global__ void read_tile_dir_kernel(const uint64_t *read_off, uint64_t n_reads, int w, uint64_t n_tiles, uint64_t *tile_first_read)
{
    uint64_t tile = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (tile >= n_tiles) return;
    long long g0 = (long long)tile * tile_cap(w, READ_TILE_THREADS) - w - tile_pad(w);
    uint64_t lo = 0, hi = n_reads + 1;                               // search over read_off[0 .. n_reads]
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if ((long long)read_off[mid] < g0) lo = mid + 1; else hi = mid;
    }
    tile_first_read[tile] = lo;
}

// ------------------------------------------------------------------ launchers

cudaError_t launch_read_tile_dir(const uint64_t *read_off, uint64_t n_reads, int w, uint64_t n_tiles, uint64_t *out, cudaStream_t st)
{
    if (!n_tiles) return cudaSuccess;
    read_tile_dir_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, st>>>(read_off, n_reads, w, n_tiles, out);
    return cudaGetLastError();
}

cudaError_t launch_read_sketch(const ReadSketchArgs &A, uint64_t n_tiles, cudaStream_t st)   // tiles [A.tile0, A.tile0 + n_tiles)
{
    if (!n_tiles) return cudaSuccess;
    size_t smem = (size_t)A.layout.bytes;
    cudaError_t e = cudaFuncSetAttribute(read_sketch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    read_sketch_kernel<<<(unsigned)n_tiles, NT, smem, st>>>(A);
    return cudaGetLastError();
}

int read_tile_windows(int w) { return tile_cap(w, READ_TILE_THREADS); }
TileLayout read_tile_layout(int k, int w) { return make_layout(k, w, false); }

}  // namespace phi
