// particles.cu -- ParticleGroup storage (src/particle_group.jl): device-resident fp64 SoA,
// conversion from/to the reference's (D+V+W) x N record layout, periodic cell sort and
// device-side synthetic loads.
#include "objects.cuh"

namespace gempic {

// ---- AoS <-> SoA -------------------------------------------------------------------------
// Staged through a bounded device buffer; the transpose runs through shared memory so that
// both the record side and the row side are accessed with full 128 B lines.
constexpr int kTile = 256;  // particles per block-iteration

__global__ void k_deinterleave(const double *__restrict__ aos, double *__restrict__ soa, size_t stride, int rows,
                               int64_t first, int64_t count)
{
    extern __shared__ double tile[];  // kTile * rows
    for (int64_t base = (int64_t)blockIdx.x * kTile; base < count; base += (int64_t)gridDim.x * kTile) {
        const int np = (int)min((int64_t)kTile, count - base);
        for (int i = threadIdx.x; i < np * rows; i += blockDim.x) tile[i] = aos[base * rows + i];
        __syncthreads();
        for (int i = threadIdx.x; i < np * rows; i += blockDim.x) {
            const int r = i / np, p = i - r * np;
            soa[(size_t)r * stride + first + base + p] = tile[p * rows + r];
        }
        __syncthreads();
    }
}

__global__ void k_interleave(double *__restrict__ aos, const double *__restrict__ soa, size_t stride, int rows,
                             int64_t first, int64_t count)
{
    extern __shared__ double tile[];
    for (int64_t base = (int64_t)blockIdx.x * kTile; base < count; base += (int64_t)gridDim.x * kTile) {
        const int np = (int)min((int64_t)kTile, count - base);
        for (int i = threadIdx.x; i < np * rows; i += blockDim.x) {
            const int r = i / np, p = i - r * np;
            tile[p * rows + r] = soa[(size_t)r * stride + first + base + p];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < np * rows; i += blockDim.x) aos[base * rows + i] = tile[i];
        __syncthreads();
    }
}

static const int64_t kStageParticles = 1 << 22;  // 4 Mi records per staging chunk

void pg_upload(ParticleGroup &pg, const double *aos)
{
    particles_changed();
    Context &c = ctx();
    pg.sorted2d = false;
    const int rows = pg.rows();
    const int64_t chunk = std::min<int64_t>(pg.n, kStageParticles);
    DevBuf<double> stage((size_t)chunk * rows);
    for (int64_t first = 0; first < pg.n; first += chunk) {
        const int64_t count = std::min<int64_t>(chunk, pg.n - first);
        GP_CUDA(cudaMemcpyAsync(stage.p, aos + (size_t)first * rows, sizeof(double) * count * rows, cudaMemcpyHostToDevice,
                                c.stream));
        const int grid = (int)std::min<int64_t>((count + kTile - 1) / kTile, (int64_t)c.sm_count * 8);
        k_deinterleave<<<grid, 256, kTile * rows * sizeof(double), c.stream>>>(stage.p, pg.data.p, pg.stride, rows, first, count);
        GP_CUDA(cudaGetLastError());
        count_launch();
        GP_CUDA(cudaStreamSynchronize(c.stream));
    }
}

void pg_download(ParticleGroup &pg, double *aos)
{
    Context &c = ctx();
    const int rows = pg.rows();
    const int64_t chunk = std::min<int64_t>(pg.n, kStageParticles);
    DevBuf<double> stage((size_t)chunk * rows);
    for (int64_t first = 0; first < pg.n; first += chunk) {
        const int64_t count = std::min<int64_t>(chunk, pg.n - first);
        const int grid = (int)std::min<int64_t>((count + kTile - 1) / kTile, (int64_t)c.sm_count * 8);
        k_interleave<<<grid, 256, kTile * rows * sizeof(double), c.stream>>>(stage.p, pg.data.p, pg.stride, rows, first, count);
        GP_CUDA(cudaGetLastError());
        count_launch();
        GP_CUDA(cudaMemcpyAsync(aos + (size_t)first * rows, stage.p, sizeof(double) * count * rows, cudaMemcpyDeviceToHost,
                                c.stream));
        GP_CUDA(cudaStreamSynchronize(c.stream));
    }
}

// ---- periodic cell sort (stable counting sort) ---------------------------------------------
// Particles are split into contiguous warp-chunks. Pass 1 counts cells per warp-chunk, a
// scan in (cell-major, chunk-minor) order turns the counts into start offsets, pass 2 walks
// every chunk again in order and ranks the 32 particles of a step with __match_any_sync, so
// the sort is stable and deterministic (no atomics on the ordering path).
constexpr int kSortBlock = 128;
constexpr int kSortWarps = kSortBlock / 32;

__device__ __forceinline__ int sort_cell(double x, const Mesh1D &m)
{
    int c;
    double t;
    cell_offset(x, m, c, t);
    return wrap_index(c, m);
}

__global__ void k_sort_count(const double *__restrict__ x, int64_t n, Mesh1D m, int64_t chunk, int n_chunks,
                             int *__restrict__ counts /* [n_cells][n_chunks] */)
{
    extern __shared__ int scount[];  // kSortWarps * n_cells
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *mine = scount + warp * m.n;
    for (int wc = blockIdx.x * kSortWarps + warp; wc < n_chunks; wc += gridDim.x * kSortWarps) {
        for (int i = lane; i < m.n; i += 32) mine[i] = 0;
        __syncwarp();
        const int64_t lo = (int64_t)wc * chunk, hi = min(n, lo + chunk);
        for (int64_t i = lo + lane; i < hi; i += 32) atomicAdd(&mine[sort_cell(x[i], m)], 1);
        __syncwarp();
        for (int i = lane; i < m.n; i += 32) counts[(size_t)i * n_chunks + wc] = mine[i];
        __syncwarp();
    }
}

// exclusive scan of `total` ints by one block (total = n_cells * n_chunks, ~1e5)
__global__ void k_sort_scan(int *__restrict__ counts, int64_t total)
{
    __shared__ int64_t carry;
    __shared__ int64_t wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < total; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        int64_t v = i < total ? counts[i] : 0;
        int64_t incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int64_t s = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                int64_t t = __shfl_up_sync(0xffffffffu, s, off);
                if (lane >= off) s += t;
            }
            wsum[lane] = s;  // inclusive over warps
        }
        __syncthreads();
        const int64_t before = carry + (warp ? wsum[warp - 1] : 0) + incl - v;
        if (i < total) counts[i] = (int)before;
        __syncthreads();
        if (threadIdx.x == 0) carry += wsum[(blockDim.x >> 5) - 1];
        __syncthreads();
    }
}

// particle i of src -> slot d of dst, all rows: every row load is issued before the first store, so a thread keeps
// `rows` loads in flight instead of one (the scatter is latency bound otherwise: 2048 threads/SM x 8 B)
constexpr int kMaxRows = 8;
__device__ __forceinline__ void move_rows(const double *__restrict__ src, double *__restrict__ dst, size_t stride, int rows,
                                          int64_t i, int64_t d)
{
    double v[kMaxRows];
#pragma unroll
    for (int r = 0; r < kMaxRows; ++r)
        if (r < rows) v[r] = src[(size_t)r * stride + i];
#pragma unroll
    for (int r = 0; r < kMaxRows; ++r)
        if (r < rows) dst[(size_t)r * stride + d] = v[r];
    for (int r = kMaxRows; r < rows; ++r) dst[(size_t)r * stride + d] = src[(size_t)r * stride + i];
}

__global__ void k_sort_scatter(const double *__restrict__ src, double *__restrict__ dst, size_t stride, int rows,
                               int64_t n, Mesh1D m, int64_t chunk, int n_chunks, const int *__restrict__ offsets)
{
    extern __shared__ int sbase[];  // kSortWarps * n_cells
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *mine = sbase + warp * m.n;
    for (int wc = blockIdx.x * kSortWarps + warp; wc < n_chunks; wc += gridDim.x * kSortWarps) {
        for (int i = lane; i < m.n; i += 32) mine[i] = offsets[(size_t)i * n_chunks + wc];
        __syncwarp();
        const int64_t lo = (int64_t)wc * chunk, hi = min(n, lo + chunk);
        for (int64_t base = lo; base < hi; base += 32) {
            const int64_t i = base + lane;
            const bool active = i < hi;
            const unsigned amask = __ballot_sync(0xffffffffu, active);
            if (active) {
                const int cell = sort_cell(src[i], m);
                const unsigned peers = __match_any_sync(amask, cell);
                const int rank = __popc(peers & ((1u << lane) - 1u));
                const int64_t d = (int64_t)mine[cell] + rank;
                move_rows(src, dst, stride, rows, i, d);
                __syncwarp(amask);
                if (rank == 0) mine[cell] += __popc(peers);
            }
            __syncwarp();
        }
        __syncwarp();
    }
}

void sort_scan(int *counts, int64_t total)
{
    k_sort_scan<<<1, 1024, 0, ctx().stream>>>(counts, total);
    GP_CUDA(cudaGetLastError());
    count_launch();
}

void pg_sort_1d(ParticleGroup &pg, const Pmc1D &p)
{
    particles_changed();
    pg.sorted2d = false;
    Context &c = ctx();
    if (pg.n < 2) return;
    GP_REQUIRE(pg.n < (int64_t)2147483647, GEMPIC_EINVAL, "sort supports < 2^31 particles per GPU");
    const Mesh1D m = p.mesh(p.Lx);
    const int grid = c.sm_count * 8;
    int n_chunks = grid * kSortWarps;
    int64_t chunk = (pg.n + n_chunks - 1) / n_chunks;
    chunk = (chunk + 31) / 32 * 32;
    n_chunks = (int)((pg.n + chunk - 1) / chunk);
    DevBuf<int> counts((size_t)m.n * n_chunks);
    if (pg.sort_tmp.n < pg.data.n) pg.sort_tmp.alloc(pg.data.n);
    const size_t smem = (size_t)kSortWarps * m.n * sizeof(int);
    GP_REQUIRE(smem <= 48 * 1024, GEMPIC_EINVAL, "sort: %d cells exceed the supported 3072", m.n);
    k_sort_count<<<grid, kSortBlock, smem, c.stream>>>(pg.row(0), pg.n, m, chunk, n_chunks, counts.p);
    GP_CUDA(cudaGetLastError());
    k_sort_scan<<<1, 1024, 0, c.stream>>>(counts.p, (int64_t)m.n * n_chunks);
    GP_CUDA(cudaGetLastError());
    k_sort_scatter<<<grid, kSortBlock, smem, c.stream>>>(pg.data.p, pg.sort_tmp.p, pg.stride, pg.rows(), pg.n, m, chunk,
                                                         n_chunks, counts.p);
    GP_CUDA(cudaGetLastError());
    count_launch(3);
    GP_CUDA(cudaStreamSynchronize(c.stream));
    std::swap(pg.data.p, pg.sort_tmp.p);
    std::swap(pg.data.n, pg.sort_tmp.n);
    pg.generation++;
}

// ---- 2D cell sort (ParticleGroup{2,V}) ------------------------------------------------------------
// key = cx + cy*nx on the TwoDMaxwell mesh.  Global histogram -> exclusive scan -> scatter with one
// warp-aggregated cursor reservation per (32 particles, distinct cell).  The order inside a cell depends
// on the reservation order (only the fp64 summation order of later deposits changes with it).
struct SortMesh2 {
    double xmin[2], inv_d[2];
    int n[2];
};
__device__ __forceinline__ int sort_cell2(double x, double y, const SortMesh2 &m)
{
    int cx = __double2int_rd((x - m.xmin[0]) * m.inv_d[0]), cy = __double2int_rd((y - m.xmin[1]) * m.inv_d[1]);
    cx = min(max(cx, 0), m.n[0] - 1);   // positions are kept inside the box; clamp the rounding edge
    cy = min(max(cy, 0), m.n[1] - 1);
    return cx + cy * m.n[0];
}
// per-block histogram of a contiguous chunk: shared counters (warp-aggregated), one row of `blockhist` per block,
// non-zero entries added to the global histogram
__global__ void __launch_bounds__(256) k_sort2_hist(const double *__restrict__ x, const double *__restrict__ y, int64_t n,
                                                    int64_t per_block, SortMesh2 m, int ncell, int *__restrict__ hist,
                                                    int *__restrict__ blockhist)
{
    extern __shared__ int sh[];
    for (int i = threadIdx.x; i < ncell; i += 256) sh[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t lo = (int64_t)blockIdx.x * per_block, hi = min(n, lo + per_block);
    for (int64_t base = lo + (threadIdx.x & ~31); base < hi; base += 256) {
        const int64_t i = base + lane;
        const bool active = i < hi;
        const unsigned amask = __ballot_sync(0xffffffffu, active);
        if (active) {
            const int cell = sort_cell2(x[i], y[i], m);
            const unsigned peers = __match_any_sync(amask, cell);
            if ((peers & ((1u << lane) - 1u)) == 0) atomicAdd(&sh[cell], __popc(peers));
        }
    }
    __syncthreads();
    int *row = blockhist + (size_t)blockIdx.x * ncell;
    for (int i = threadIdx.x; i < ncell; i += 256) {
        const int c = sh[i];
        row[i] = c;
        if (c) atomicAdd(&hist[i], c);
    }
}

// scatter of the same chunk: reserve [cursor, cursor + count) per cell for this block, then place the particles
__global__ void __launch_bounds__(256) k_sort2_scatter(const double *__restrict__ src, double *__restrict__ dst, size_t stride,
                                                       int rows, int64_t n, int64_t per_block, SortMesh2 m, int ncell,
                                                       int *__restrict__ cursor, const int *__restrict__ blockhist)
{
    extern __shared__ int sh[];   // next free slot of every cell for this block
    const int *row = blockhist + (size_t)blockIdx.x * ncell;
    for (int i = threadIdx.x; i < ncell; i += 256) {
        const int c = row[i];
        sh[i] = c ? atomicAdd(&cursor[i], c) : 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t lo = (int64_t)blockIdx.x * per_block, hi = min(n, lo + per_block);
    for (int64_t base = lo + (threadIdx.x & ~31); base < hi; base += 256) {
        const int64_t i = base + lane;
        const bool active = i < hi;
        const unsigned amask = __ballot_sync(0xffffffffu, active);
        if (active) {
            const int cell = sort_cell2(src[i], src[stride + i], m);
            const unsigned peers = __match_any_sync(amask, cell);
            const int leader = __ffs(peers) - 1;
            int start = 0;
            if (lane == leader) start = atomicAdd(&sh[cell], __popc(peers));
            start = __shfl_sync(peers, start, leader);
            const int64_t d = (int64_t)start + __popc(peers & ((1u << lane) - 1u));
            move_rows(src, dst, stride, rows, i, d);
        }
    }
}

void pg_sort_2d(ParticleGroup &pg, const Maxwell2D &mx)
{
    particles_changed();
    Context &c = ctx();
    if (pg.n < 2) return;
    GP_REQUIRE(pg.D == 2, GEMPIC_EINVAL, "2D cell sort needs a D = 2 particle group");
    GP_REQUIRE(pg.n < (int64_t)2147483647, GEMPIC_EINVAL, "sort supports < 2^31 particles per GPU");
    SortMesh2 m;
    m.xmin[0] = mx.xmin; m.xmin[1] = mx.ymin;
    m.inv_d[0] = 1.0 / mx.dx; m.inv_d[1] = 1.0 / mx.dy;
    m.n[0] = mx.nx; m.n[1] = mx.ny;
    const int64_t cells = (int64_t)mx.nx * mx.ny;
    const size_t smem = (size_t)cells * sizeof(int);
    GP_REQUIRE(smem <= 200 * 1024, GEMPIC_EINVAL, "2D cell sort supports up to 51200 cells (%lld requested)", (long long)cells);
    const int grid = (int)std::min<int64_t>((int64_t)c.sm_count * 8, (pg.n + 255) / 256);
    int64_t per_block = (pg.n + grid - 1) / grid;
    per_block = (per_block + 255) / 256 * 256;
    const size_t need = (size_t)cells * ((size_t)grid + 1);
    if (pg.sort_keys.n < need) pg.sort_keys.alloc(need);
    if (pg.sort_tmp.n < pg.data.n) pg.sort_tmp.alloc(pg.data.n);
    int *hist = pg.sort_keys.p, *blockhist = pg.sort_keys.p + cells;
    if (smem > 48 * 1024) {
        ensure_func_smem((const void *)k_sort2_hist, 200 * 1024);
        ensure_func_smem((const void *)k_sort2_scatter, 200 * 1024);
    }
    GP_CUDA(cudaMemsetAsync(hist, 0, sizeof(int) * cells, c.stream));
    profile_begin("cell sort 2d");
    k_sort2_hist<<<grid, 256, smem, c.stream>>>(pg.row(0), pg.row(1), pg.n, per_block, m, (int)cells, hist, blockhist);
    GP_CUDA(cudaGetLastError());
    k_sort_scan<<<1, 1024, 0, c.stream>>>(hist, cells);
    GP_CUDA(cudaGetLastError());
    k_sort2_scatter<<<grid, 256, smem, c.stream>>>(pg.data.p, pg.sort_tmp.p, pg.stride, pg.rows(), pg.n, per_block, m,
                                                   (int)cells, hist, blockhist);
    GP_CUDA(cudaGetLastError());
    profile_end("cell sort 2d");
    count_launch(3);
    GP_CUDA(cudaStreamSynchronize(c.stream));
    std::swap(pg.data.p, pg.sort_tmp.p);
    std::swap(pg.data.n, pg.sort_tmp.n);
    pg.generation++;
    pg.sorted2d = true;
}

// ---- synthetic loads -------------------------------------------------------------------------
// Counter-based generator: splitmix64 of (seed, global particle index, stream) -> reproducible
// for any sharding of the index range.  Statistical (not bitwise) counterpart of
// src/particle_sampling.jl; parity tests exchange particle arrays instead.
__device__ __forceinline__ uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(uint64_t seed, uint64_t idx, uint64_t stream)
{
    const uint64_t h = splitmix64(splitmix64(seed ^ (stream * 0xD1B54A32D192ED03ull)) + idx);
    return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);  // (0,1)
}

struct SampleParams {
    int kind;
    double xmin, L, alpha, k;
    double sigma[3];
    uint64_t seed;
    int64_t first;
};

__global__ void k_sample(double *__restrict__ data, size_t stride, int D, int V, int W, int64_t n, SampleParams s)
{
    const double two_pi = 6.283185307179586476925286766559;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t g = (uint64_t)(s.first + i);
        for (int d = 0; d < D; ++d) {
            const double u = u01(s.seed, g, 10 + d);
            double x = u * s.L;
            if (s.kind == 1 && d == 0 && s.alpha != 0.0) {
                // Landau: solve x + (alpha/k) sin(k x) = u L by Newton (particle_sampling.jl:236-245)
                for (int it = 0; it < 50; ++it) {
                    const double f = x + s.alpha / s.k * sin(s.k * x) - u * s.L;
                    const double fp = 1.0 + s.alpha * cos(s.k * x);
                    const double dx = f / fp;
                    x -= dx;
                    if (fabs(dx) < 1e-14 * s.L) break;
                }
            }
            if (x < 0.0) x += s.L;
            if (x >= s.L) x -= s.L;
            data[(size_t)d * stride + i] = s.xmin + x;
        }
        for (int v = 0; v < V; v += 2) {  // Box-Muller pairs
            const double u1 = u01(s.seed, g, 20 + v), u2 = u01(s.seed, g, 21 + v);
            const double r = sqrt(-2.0 * log(u1));
            data[(size_t)(D + v) * stride + i] = s.sigma[v] * r * cos(two_pi * u2);
            if (v + 1 < V) data[(size_t)(D + v + 1) * stride + i] = s.sigma[v + 1] * r * sin(two_pi * u2);
        }
        double wgt = s.L;
        for (int d = 1; d < D; ++d) wgt *= s.L;
        for (int w = 0; w < W; ++w) data[(size_t)(D + V + w) * stride + i] = wgt;  // w = Lx (particle_sampling.jl:308)
    }
}

void pg_sample(ParticleGroup &pg, int kind, double xmin, double L, double alpha, double k, const double *sigma,
               uint64_t seed, int64_t first_index)
{
    particles_changed();
    Context &c = ctx();
    pg.sorted2d = false;
    GP_REQUIRE(kind == 0 || kind == 1, GEMPIC_EINVAL, "unknown sample kind %d", kind);
    GP_REQUIRE(pg.V <= 3, GEMPIC_EINVAL, "at most 3 velocity dimensions");
    SampleParams s{};
    s.kind = kind; s.xmin = xmin; s.L = L; s.alpha = alpha; s.k = k; s.seed = seed; s.first = first_index;
    for (int v = 0; v < pg.V; ++v) s.sigma[v] = sigma ? sigma[v] : 1.0;
    k_sample<<<c.sm_count * 8, 256, 0, c.stream>>>(pg.data.p, pg.stride, pg.D, pg.V, pg.W, pg.n, s);
    GP_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace gempic
