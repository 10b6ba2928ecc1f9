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
//   k_grid_bounds   largest box side -> cell size G (two overlapping boxes have centres in the same or adjacent cells)
//   k_grid_count / k_grid_scan / k_grid_fill     counting sort of the candidates by cell
//   k_grid_dominators   every candidate walks the 3x3 cells around it and records its dominators (fixed capacity per box)
//   k_grid_round    one round of decisions; the host reads the number of undecided boxes every few rounds
// A box with more dominators than the capacity, or a chain longer than the round limit, makes the caller fall back to
// the general pipeline (PostProc::run) - exactness never depends on the fast path.
#include "postproc.cuh"
#include "tiles.cuh"

#include <algorithm>

namespace y3 {

static constexpr int GRID_DOM_CAP = 32;       // dominators recorded per box
static constexpr int GRID_MAX_ROUNDS = 96;

struct GridCtrl {
    int n_cand;           // candidates
    int cell;             // cell size in pixels
    int gx, gy;           // grid dimensions
    int overflow;         // a box had more than GRID_DOM_CAP dominators
    int undecided;        // boxes still undecided after the last round
    float max_side;
    int pad_;
};

__device__ __forceinline__ bool higher_priority(float sj, int rj, float si, int ri) { return sj > si || (sj == si && rj < ri); }

__global__ void __launch_bounds__(256)
k_grid_bounds(const float4* __restrict__ box, const uint8_t* __restrict__ cand, int64_t n, GridCtrl* __restrict__ G) {
    float m = 0.f;
    int c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (cand[i]) {
            const float4 b = box[i];
            m = fmaxf(m, fmaxf(b.z - b.x, b.w - b.y));
            ++c;
        }
    for (int o = 16; o; o >>= 1) { m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o)); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if ((threadIdx.x & 31) == 0) {
        if (c) atomicAdd(&G->n_cand, c);
        atomicMax(reinterpret_cast<int*>(&G->max_side), __float_as_int(m));       // m >= 0: int order == float order
    }
}

// one thread: cell size and grid dimensions (bounded so that the cell table stays small)
__global__ void k_grid_setup(GridCtrl* __restrict__ G, long long img_w, long long img_h, int max_cells_side) {
    float side = G->max_side;
    if (!(side >= 1.f)) side = 1.f;                                 // NaN / degenerate: any positive cell size works
    long long cell = (long long)ceilf(side) + 1;
    const long long need = (max(img_w, img_h) + max_cells_side - 1) / max_cells_side;
    cell = max(cell, max(need, 16ll));
    G->cell = (int)min(cell, 1ll << 30);
    G->gx = (int)((img_w + G->cell - 1) / G->cell) + 1;
    G->gy = (int)((img_h + G->cell - 1) / G->cell) + 1;
}

__device__ __forceinline__ int cell_of(const float4 b, const GridCtrl& G) {
    // centre of the box, clamped into the grid (coordinates are clamped pixel corners, NaN goes to cell 0)
    const float cx = 0.5f * (b.x + b.z), cy = 0.5f * (b.y + b.w);
    int ix = (int)(cx / (float)G.cell), iy = (int)(cy / (float)G.cell);
    ix = min(max(ix, 0), G.gx - 1);
    iy = min(max(iy, 0), G.gy - 1);
    return iy * G.gx + ix;
}

__global__ void __launch_bounds__(256)
k_grid_count(const float4* __restrict__ box, const uint8_t* __restrict__ cand, int64_t n, const GridCtrl* __restrict__ Gp,
             int* __restrict__ cell_cnt, int* __restrict__ slot) {
    const GridCtrl G = *Gp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (cand[i]) slot[i] = atomicAdd(&cell_cnt[cell_of(box[i], G)], 1);
}

__global__ void __launch_bounds__(1024)
k_grid_scan(int* __restrict__ cell_cnt, int* __restrict__ cell_off, const GridCtrl* __restrict__ Gp) {
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
    if (threadIdx.x == 1023) cell_off[n] = (int)s_part[1023];
}

// members[cell_off[cell] + slot] = row;  state: 0 undecided, 1 kept, 2 dead
__global__ void __launch_bounds__(256)
k_grid_fill(const float4* __restrict__ box, const uint8_t* __restrict__ cand, int64_t n, const GridCtrl* __restrict__ Gp,
            const int* __restrict__ cell_off, const int* __restrict__ slot, int* __restrict__ members, uint8_t* __restrict__ state) {
    const GridCtrl G = *Gp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        state[i] = 0;
        if (cand[i]) members[cell_off[cell_of(box[i], G)] + slot[i]] = (int)i;
    }
}

__global__ void __launch_bounds__(128)
k_grid_dominators(const float4* __restrict__ box, const float* __restrict__ score, const int32_t* __restrict__ label, GridCtrl* __restrict__ Gp,
                  const int* __restrict__ cell_off, const int* __restrict__ members, float thr, int* __restrict__ dom,
                  int* __restrict__ n_dom) {
    const GridCtrl G = *Gp;
    const int m = G.n_cand;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < m; p += gridDim.x * blockDim.x) {
        const int i = members[p];
        const float4 bi = box[i];
        const float ai = box_area_exact(bi);
        const float si = score[i];
        const int li = label[i];
        const int c = cell_of(bi, G);
        const int cy = c / G.gx, cx = c - cy * G.gx;
        int nd = 0;
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = cy + dy;
            if (y < 0 || y >= G.gy) continue;
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, G.gx - 1);
            const int q0 = cell_off[y * G.gx + x0], q1 = cell_off[y * G.gx + x1 + 1];     // the three cells of a row are contiguous
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
    Y3_CUDA(cudaMemsetAsync(cell_cnt, 0, (size_t)(max_side_cells + 2) * (max_side_cells + 2) * 4, st));
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
    GridCtrl* hG = grid_host.as<GridCtrl>();
    bool done = false;
    for (int round = 0; round < GRID_MAX_ROUNDS && !done; round += 4) {
        for (int r = 0; r < 4; ++r) {
            if (r == 3) { k_grid_reset_undecided<<<1, 1, 0, st>>>(G); Y3_LAUNCHED(ctx); }
            k_grid_round<<<blocks, 256, 0, st>>>(grid_members.as<int>(), grid_dom.as<int>(), grid_ndom.as<int>(), G,
                                                 grid_state.as<uint8_t>(), 0);
            Y3_LAUNCHED(ctx);
        }
        Y3_CUDA(cudaMemcpyAsync(hG, G, sizeof(GridCtrl), cudaMemcpyDeviceToHost, st));
        Y3_CUDA(cudaStreamSynchronize(st));
        if (hG->overflow) return false;
        done = hG->undecided == 0;
    }
    if (!done) return false;
    k_grid_keepmask<<<ceil_div(n, 256), 256, 0, st>>>(cand, grid_state.as<uint8_t>(), n, keepm);
    Y3_LAUNCHED(ctx);
    return true;
}

}  // namespace y3
