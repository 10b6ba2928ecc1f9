// nms_grid.cu - exact greedy per-class NMS for a SPARSE set of boxes spread over a large image (the cross-seam stage:
// tens of thousands of seam candidates, each overlapping a handful of others), as a parallel fixed-point iteration.
//
// Greedy NMS keeps box i iff no KEPT box of the same class with higher priority (score descending, then row ascending -
// the tie rule of nms.cu) suppresses it (IoU > thr, exact fp32 arithmetic of bbox_utils.compute_iou).  That is the
// lexicographically-first maximal independent set of the "suppresses" graph, and it can be evaluated in rounds:
//   a box whose higher-priority suppressors ("dominators") are all DEAD is KEPT; a box with a KEPT dominator is DEAD.
// Decisions are final and only ever use final decisions, so the result is the serial one regardless of scheduling.  The
// number of rounds is the longest dominator chain - a few tens for boxes that were already NMS-ed inside their tiles.
//
//   k_grid_bounds   histogram of the box sides -> cell size G = the power of two that holds 95 % of the boxes (two
//                   overlapping boxes no larger than a cell have centres in the same or adjacent cells); the few larger
//                   boxes ("big": clamped giants, outliers) go to a separate list instead of blowing the cell size up
//   k_grid_count / k_grid_scan / k_grid_fill     counting sort of the normal candidates by cell, big ones appended behind
//   k_grid_dominators   a normal candidate walks the 3x3 cells around it plus the big list, a big one walks everything;
//                   each records its dominators (fixed capacity per box)
//   k_grid_round    one round of decisions; the host reads the number of undecided boxes every few rounds
// A box with more dominators than the capacity, or a chain longer than the round limit, makes the caller fall back to
// the general pipeline (PostProc::run) - exactness never depends on the fast path.
#include "postproc.cuh"
#include "tiles.cuh"

#include <algorithm>
#include <stdlib.h>

namespace y3 {

static constexpr int GRID_DOM_CAP = 32;       // dominators recorded per box
static constexpr int GRID_MAX_ROUNDS = 96;

static constexpr int GRID_MAX_BIG = 4096;     // boxes larger than a cell that the fast path accepts

struct GridCtrl {
    int n_cand;           // candidates
    int cell;             // cell size in pixels
    int gx, gy;           // grid dimensions
    int overflow;         // a box had more than GRID_DOM_CAP dominators, or there are too many big boxes
    int undecided;        // boxes still undecided after the last round
    int n_big;            // candidates with a side larger than the cell
    int n_norm;           // the others (cell-sorted part of `members`)
    int hist[32];         // candidates by ceil(log2(largest side))
};

__device__ __forceinline__ bool higher_priority(float sj, int rj, float si, int ri) { return sj > si || (sj == si && rj < ri); }

__device__ __forceinline__ float box_side(const float4 b) { return fmaxf(b.z - b.x, b.w - b.y); }
__device__ __forceinline__ int side_bucket(float side) {           // smallest k with side <= 2^k (0 for side <= 1, NaN -> 31)
    if (!(side <= 1.0e9f)) return 31;
    int k = 0;
    while (k < 31 && (float)(1u << k) < side) ++k;
    return k;
}

__global__ void __launch_bounds__(256)
k_grid_bounds(const float4* __restrict__ box, const uint8_t* __restrict__ cand, int64_t n, GridCtrl* __restrict__ G) {
    __shared__ int s_hist[32];
    if (threadIdx.x < 32) s_hist[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (cand[i]) atomicAdd(&s_hist[side_bucket(box_side(box[i]))], 1);
    __syncthreads();
    if (threadIdx.x < 32 && s_hist[threadIdx.x]) atomicAdd(&G->hist[threadIdx.x], s_hist[threadIdx.x]);
}

// one thread: cell size = the power of two that holds 95 % of the candidates (at least 64 px, and large enough that the
// cell table stays within max_cells_side^2)
__global__ void k_grid_setup(GridCtrl* __restrict__ G, long long img_w, long long img_h, int max_cells_side) {
    long long total = 0;
    for (int k = 0; k < 32; ++k) total += G->hist[k];
    G->n_cand = (int)total;
    long long acc = 0;
    int kb = 6;
    for (int k = 0; k < 31; ++k) {
        acc += G->hist[k];
        if (k >= 6 && acc * 100 >= total * 95) { kb = k; break; }
        kb = k + 1;
    }
    const long long need = (max(img_w, img_h) + max_cells_side - 1) / max_cells_side;
    kb = min(kb, 10);                                               // cells of at most 1024 px (unless the table bound asks for more):
    long long cell = (1ll << kb) + 1;                               // whatever is larger is a "big" box
    cell = max(cell, need);
    long long big = 0;
    for (int k = 0; k < 32; ++k) if ((1ll << k) > cell - 1) big += G->hist[k];     // upper bound of the big list
    if (big > GRID_MAX_BIG) G->overflow = 1;                        // heavy-tailed sizes: not a sparse problem
    G->cell = (int)min(cell, 1ll << 30);
    G->gx = (int)((img_w + G->cell - 1) / G->cell) + 1;
    G->gy = (int)((img_h + G->cell - 1) / G->cell) + 1;
}

__device__ __forceinline__ bool is_big(const float4 b, const GridCtrl& G) { return !(box_side(b) <= (float)(G.cell - 1)); }

__device__ __forceinline__ int cell_of(const float4 b, const GridCtrl& G) {
    // centre of the box, clamped into the grid (coordinates are clamped pixel corners, NaN goes to cell 0)
    const float cx = 0.5f * (b.x + b.z), cy = 0.5f * (b.y + b.w);
    int ix = (int)(cx / (float)G.cell), iy = (int)(cy / (float)G.cell);
    ix = min(max(ix, 0), G.gx - 1);
    iy = min(max(iy, 0), G.gy - 1);
    return iy * G.gx + ix;
}

__global__ void __launch_bounds__(256)
k_grid_count(const float4* __restrict__ box, const uint8_t* __restrict__ cand, int64_t n, GridCtrl* __restrict__ Gp,
             int* __restrict__ cell_cnt, int* __restrict__ slot) {
    const GridCtrl G = *Gp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (cand[i]) {
            const float4 b = box[i];
            if (is_big(b, G)) slot[i] = -1 - atomicAdd(&Gp->n_big, 1);            // negative: index in the big list
            else slot[i] = atomicAdd(&cell_cnt[cell_of(b, G)], 1);
        }
}

__global__ void __launch_bounds__(1024)
k_grid_scan(int* __restrict__ cell_cnt, int* __restrict__ cell_off, GridCtrl* __restrict__ Gp) {
    __shared__ long long s_part[1024];
    const int n = Gp->gx * Gp->gy;
    const int per = (n + 1023) / 1024;
    const int b0 = threadIdx.x * per;
    long long sum = 0;
    for (int i = 0; i < per; ++i) if (b0 + i < n) sum += cell_cnt[b0 + i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const long long v = (threadIdx.x >= o) ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    long long run = s_part[threadIdx.x] - sum;
    for (int i = 0; i < per; ++i)
        if (b0 + i < n) { cell_off[b0 + i] = (int)run; run += cell_cnt[b0 + i]; }
    if (threadIdx.x == 1023) {
        cell_off[n] = (int)s_part[1023];
        Gp->n_norm = (int)s_part[1023];
        if (Gp->n_big > GRID_MAX_BIG) Gp->overflow = 1;
    }
}

// members[cell_off[cell] + slot] = row;  state: 0 undecided, 1 kept, 2 dead
__global__ void __launch_bounds__(256)
k_grid_fill(const float4* __restrict__ box, const uint8_t* __restrict__ cand, int64_t n, const GridCtrl* __restrict__ Gp,
            const int* __restrict__ cell_off, const int* __restrict__ slot, int* __restrict__ members, uint8_t* __restrict__ state) {
    const GridCtrl G = *Gp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        state[i] = 0;
        if (!cand[i]) continue;
        const int sl = slot[i];
        if (sl < 0) members[G.n_norm + (-1 - sl)] = (int)i;                 // big boxes behind the cell-sorted ones
        else members[cell_off[cell_of(box[i], G)] + sl] = (int)i;
    }
}

__global__ void __launch_bounds__(128)
k_grid_dominators(const float4* __restrict__ box, const float* __restrict__ score, const int32_t* __restrict__ label, GridCtrl* __restrict__ Gp,
                  const int* __restrict__ cell_off, const int* __restrict__ members, float thr, int* __restrict__ dom,
                  int* __restrict__ n_dom) {
    const GridCtrl G = *Gp;
    const int m = G.n_cand;
    if (G.overflow) return;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < m; p += gridDim.x * blockDim.x) {
        const int i = members[p];
        const float4 bi = box[i];
        const float ai = box_area_exact(bi);
        const float si = score[i];
        const int li = label[i];
        int nd = 0;
        auto visit = [&](int q0, int q1) {
            for (int q = q0; q < q1; ++q) {
                const int j = members[q];
                if (j == i || label[j] != li) continue;
                const float sj = score[j];
                if (!higher_priority(sj, j, si, i)) continue;
                const float4 bj = box[j];
                // the picked (higher-priority) box first: the operand order of single_class_nms's IoU row
                if (suppresses_exact(bj, box_area_exact(bj), bi, ai, thr)) {
                    if (nd < GRID_DOM_CAP) dom[(size_t)p * GRID_DOM_CAP + nd] = j;
                    ++nd;
                }
            }
        };
        if (p >= G.n_norm) {
            visit(0, m);                                          // a big box: everything
        } else {
            const int c = cell_of(bi, G);
            const int cy = c / G.gx, cx = c - cy * G.gx;
            for (int dy = -1; dy <= 1; ++dy) {
                const int y = cy + dy;
                if (y < 0 || y >= G.gy) continue;
                const int x0 = max(cx - 1, 0), x1 = min(cx + 1, G.gx - 1);
                visit(cell_off[y * G.gx + x0], cell_off[y * G.gx + x1 + 1]);     // the three cells of a row are contiguous
            }
            visit(G.n_norm, m);                                   // plus the big boxes
        }
        n_dom[p] = nd;
        if (nd > GRID_DOM_CAP) Gp->overflow = 1;
    }
}

__global__ void __launch_bounds__(256)
k_grid_round(const int* __restrict__ members, const int* __restrict__ dom, const int* __restrict__ n_dom, GridCtrl* __restrict__ Gp,
             volatile uint8_t* __restrict__ state, int reset_counter) {
    const int m = Gp->n_cand;
    int undecided = 0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < m; p += gridDim.x * blockDim.x) {
        const int i = members[p];
        if (state[i]) continue;
        const int nd = min(n_dom[p], GRID_DOM_CAP);
        bool any_kept = false, any_open = false;
        for (int k = 0; k < nd; ++k) {
            const uint8_t s = state[dom[(size_t)p * GRID_DOM_CAP + k]];
            any_kept |= (s == 1);
            any_open |= (s == 0);
        }
        if (any_kept) state[i] = 2;
        else if (!any_open) state[i] = 1;
        else ++undecided;
    }
    (void)reset_counter;
    for (int o = 16; o; o >>= 1) undecided += __shfl_xor_sync(0xffffffffu, undecided, o);
    if ((threadIdx.x & 31) == 0 && undecided) atomicAdd(&Gp->undecided, undecided);
}

__global__ void k_grid_reset_undecided(GridCtrl* __restrict__ G) { G->undecided = 0; }

__global__ void __launch_bounds__(256)
k_grid_keepmask_all(int64_t n, uint8_t* __restrict__ keepm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keepm[i] = 1;
}

__global__ void __launch_bounds__(256)
k_grid_keepmask(const uint8_t* __restrict__ cand, const uint8_t* __restrict__ state, int64_t n, uint8_t* __restrict__ keepm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keepm[i] = (!cand[i] || state[i] == 1) ? 1 : 0;
}

// -> true: keepm[] holds the exact result.  false: the fast path does not apply (dominator overflow / long chain).
bool Tiler::cross_seam_grid(const float4* box, const float* score, const int32_t* label, const uint8_t* cand, int64_t n,
                            const StitchArgs& S, float iou_thr, uint8_t* keepm) {
    cudaStream_t st = ctx->stream;
    if (n >= (1ll << 31)) return false;
    const int max_side_cells = 1024;
    grid_ctrl.reserve(sizeof(GridCtrl));
    grid_cells.reserve((size_t)(max_side_cells + 2) * (max_side_cells + 2) * 4 * 2 + 8);
    grid_slot.reserve((size_t)n * 4); grid_members.reserve((size_t)n * 4); grid_state.reserve((size_t)n);
    grid_dom.reserve((size_t)n * GRID_DOM_CAP * 4); grid_ndom.reserve((size_t)n * 4);
    grid_host.reserve(sizeof(GridCtrl));
    GridCtrl* G = grid_ctrl.as<GridCtrl>();
    int* cell_cnt = grid_cells.as<int>();
    int* cell_off = cell_cnt + (size_t)(max_side_cells + 2) * (max_side_cells + 2);
    Y3_CUDA(cudaMemsetAsync(G, 0, sizeof(GridCtrl), st));
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8);
    k_grid_bounds<<<blocks, 256, 0, st>>>(box, cand, n, G);
    Y3_LAUNCHED(ctx);
    k_grid_setup<<<1, 1, 0, st>>>(G, (long long)S.img_w, (long long)S.img_h, max_side_cells);
    Y3_LAUNCHED(ctx);
    GridCtrl* hG = grid_host.as<GridCtrl>();
    Y3_CUDA(cudaMemcpyAsync(hG, G, sizeof(GridCtrl), cudaMemcpyDeviceToHost, st));
    Y3_CUDA(cudaStreamSynchronize(st));
    static const bool dbg = getenv("Y3_DEBUG_TIMING") != nullptr;
    if (hG->overflow || hG->n_cand == 0) {
        if (dbg) fprintf(stderr, "y3: cross-seam grid: %d candidates, cell %d px: too many oversized boxes for the sparse path\n", hG->n_cand, hG->cell);
        if (hG->n_cand == 0) { k_grid_keepmask_all<<<ceil_div(n, 256), 256, 0, st>>>(n, keepm); Y3_LAUNCHED(ctx); return true; }
        return false;
    }
    const size_t n_cells = (size_t)hG->gx * hG->gy;
    Y3_CUDA(cudaMemsetAsync(cell_cnt, 0, n_cells * 4, st));
    k_grid_count<<<blocks, 256, 0, st>>>(box, cand, n, G, cell_cnt, grid_slot.as<int>());
    Y3_LAUNCHED(ctx);
    k_grid_scan<<<1, 1024, 0, st>>>(cell_cnt, cell_off, G);
    Y3_LAUNCHED(ctx);
    k_grid_fill<<<blocks, 256, 0, st>>>(box, cand, n, G, cell_off, grid_slot.as<int>(), grid_members.as<int>(), grid_state.as<uint8_t>());
    Y3_LAUNCHED(ctx);
    const int dblocks = (int)std::min<int64_t>((n + 127) / 128, (int64_t)ctx->sm_count * 16);
    k_grid_dominators<<<dblocks, 128, 0, st>>>(box, score, label, G, cell_off, grid_members.as<int>(), iou_thr, grid_dom.as<int>(),
                                               grid_ndom.as<int>());
    Y3_LAUNCHED(ctx);
    bool done = false;
    int rounds_run = 0;
    for (int round = 0; round < GRID_MAX_ROUNDS && !done; round += 4) {
        rounds_run = round + 4;
        for (int r = 0; r < 4; ++r) {
            if (r == 3) { k_grid_reset_undecided<<<1, 1, 0, st>>>(G); Y3_LAUNCHED(ctx); }
            k_grid_round<<<blocks, 256, 0, st>>>(grid_members.as<int>(), grid_dom.as<int>(), grid_ndom.as<int>(), G,
                                                 grid_state.as<uint8_t>(), 0);
            Y3_LAUNCHED(ctx);
        }
        Y3_CUDA(cudaMemcpyAsync(hG, G, sizeof(GridCtrl), cudaMemcpyDeviceToHost, st));
        Y3_CUDA(cudaStreamSynchronize(st));
        if (dbg) {
            fprintf(stderr, "y3: cross-seam grid: %d candidates (%d big), cell %d px, after %d rounds %d undecided, overflow %d; sides by log2:", hG->n_cand,
                    hG->n_big, hG->cell, rounds_run, hG->undecided, hG->overflow);
            for (int k = 0; k < 32; ++k) if (hG->hist[k]) fprintf(stderr, " 2^%d:%d", k, hG->hist[k]);
            fprintf(stderr, "\n");
        }
        if (hG->overflow) return false;
        done = hG->undecided == 0;
    }
    if (!done) return false;
    k_grid_keepmask<<<ceil_div(n, 256), 256, 0, st>>>(cand, grid_state.as<uint8_t>(), n, keepm);
    Y3_LAUNCHED(ctx);
    return true;
}

}  // namespace y3
