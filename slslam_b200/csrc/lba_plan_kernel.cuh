// The per-solve "symbolic" plan of an LBA window, built ON DEVICE from the caller's raw arrays (camera_index,
// line_index, fixed_index, observations): what build_plan() in lba_host.cu does on the host -- sticky constants per
// block (reference src/lba_problem.cpp:88-91), observations grouped by line (stable), lines partitioned over the CTAs
// of the window's group, whole lines packed into 32-lane tiles, conflict-free accumulator rounds, the camera-pair list
// of the Schur blocks -- with bit-identical output (tests compare the two), so the solve kernel cannot tell which
// planner ran.  With it the host only copies the caller's arrays into pinned memory: no sort, no per-observation loop.
// One CTA of PLAN_NT threads per window; every array it produces is integer / byte work (HBM- and latency-bound).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lba_kernel.cuh"

namespace slslam {

constexpr int PLAN_NT = 1024;
constexpr int PLAN_NW = PLAN_NT / 32;

// plan status flags (PlanInfo::error)
constexpr int PLAN_OK = 0, PLAN_ERR_INDEX = 1, PLAN_ERR_LIMIT = 2, PLAN_DUPLICATE_CAMERA = 4, PLAN_ERR_CAPACITY = 8;

struct PlanInfo {
  int error, Cf, max_lines_cta, max_slots_cta, max_items_cta, nslots, nitems, has_unobserved;
  int phase_cycles[8];   // diagnostics: SM cycles of thread 0 up to the end of phases A, C, D, F, G, H, I, J
};

struct PlanIn {
  int C, L, N, CS;
  int slot_cap, item_cap;
  // caller's arrays as uploaded
  const int* cam_idx;        // [N]
  const int* line_idx;       // [N]
  const int* fixed;          // [2N]
  const double* obs_raw;     // [N][8]
  // plan outputs (read by lba_solve_kernel through the WinHdr)
  double* obs;               // [slot_cap][8] slot order
  int2* meta;                // [slot_cap]
  int* line_gid;             // [L]
  uint32_t* items;           // [item_cap]
  int* key_off;              // [CS][nkeys + 1]
  WinHdr* hdr;               // Cf, n, nkeys, vlen, vpad, cta_slot_off, cta_line_off, cam_free are written here
  PlanInfo* info;
  // scratch in global memory (L2-resident at these sizes)
  int* line_cnt;             // [L]
  int* line_start;           // [L + 1]
  int* fill;                 // [L]
  int* lconst;               // [L]
  int* order;                // [N] observation indices grouped by line, ascending inside a line
  int* slot_line;            // [slot_cap] device line of the slot or -1; on exit the caller's observation index or -1
};

// Block-wide exclusive scan of v (one value per thread); returns the exclusive prefix, *total gets the block total.
// wsum: PLAN_NW + 1 ints of shared memory.
__device__ __forceinline__ int plan_block_exscan(int v, int* wsum, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < PLAN_NW ? wsum[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < PLAN_NW) wsum[lane] = winc - w;
    if (lane == 31) wsum[PLAN_NW] = winc;
  }
  __syncthreads();
  const int ex = wsum[warp] + inc - v;
  *total = wsum[PLAN_NW];
  __syncthreads();
  return ex;
}

__global__ void __launch_bounds__(PLAN_NT, 1) lba_plan_kernel(const PlanIn* __restrict__ ins) {
  const PlanIn& p = ins[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C, L = p.L, N = p.N, CS = p.CS;
  __shared__ unsigned s_cam_used, s_cam_const;
  __shared__ int s_err, s_unobs, s_Cf, s_nkeys, s_nd, s_total_slots, s_max_lines, s_max_slots, s_max_items, s_total_items;
  __shared__ int s_cam_free[MAX_CAMS];
  __shared__ int s_cta_line_off[MAX_G + 1], s_cta_slot_off[MAX_G + 1];
  __shared__ int s_wsum[PLAN_NW + 1];
  // per device line, in dynamic shared memory (13 + min(C, 24) bytes per line of the window): the sequential and pointer-chasing
  // steps below would otherwise pay an L2 round trip per line
  extern __shared__ __align__(16) unsigned char plan_dyn[];
  int* s_line = reinterpret_cast<int*>(plan_dyn);                  // first_slot (CTA-local) | seg_start << 16 | cta << 22
  unsigned* s_mask = reinterpret_cast<unsigned*>(plan_dyn) + L;    // reduced cameras observing the line
  int* s_dstart = reinterpret_cast<int*>(plan_dyn) + 2 * (size_t)L;   // first entry of the line in `order`
  unsigned char* s_cnt = plan_dyn + 12 * (size_t)L;                // observations of the line (<= 32)
  unsigned char* s_pos = plan_dyn + 13 * (size_t)L;                // [L][pstride] position in the line of reduced camera cf
  const int pstride = min(C, (int)MAX_FREE_CAMS);

  __shared__ int s_ph[8];
  const long long t_start = clock64();
#define PLAN_PHASE(i) { if (tid == 0) s_ph[i] = (int)(clock64() - t_start); }
  if (tid == 0) {
    s_cam_used = 0u; s_cam_const = 0u; s_err = PLAN_OK; s_unobs = 0; s_max_lines = 1; s_max_slots = 32; s_max_items = 0;
    for (int k = 0; k < 8; ++k) s_ph[k] = 0;
  }
  for (int l = tid; l < L; l += PLAN_NT) { p.line_cnt[l] = 0; p.lconst[l] = 0; }
  __syncthreads();

  // ---- A: counts per line, sticky constants, index validation (camera sets combined per warp before they touch the
  // two shared words, which every thread of the CTA would otherwise hammer) ----
  for (int base = 0; base < N; base += PLAN_NT) {
    const int i = base + tid;
    unsigned used = 0u, cst = 0u, bad = 0u;
    if (i < N) {
      const int c = p.cam_idx[i], l = p.line_idx[i];
      if (c < 0 || c >= C || l < 0 || l >= L) {
        bad = 1u;
      } else {
        atomicAdd(&p.line_cnt[l], 1);
        used = 1u << c;
        const int2 fx = reinterpret_cast<const int2*>(p.fixed)[i];
        if (fx.x) cst = 1u << c;
        if (fx.y) p.lconst[l] = 1;
      }
    }
    used = __reduce_or_sync(0xffffffffu, used); cst = __reduce_or_sync(0xffffffffu, cst); bad = __reduce_or_sync(0xffffffffu, bad);
    if (lane == 0) {
      if (used) atomicOr(&s_cam_used, used);
      if (cst) atomicOr(&s_cam_const, cst);
      if (bad) atomicOr(&s_err, PLAN_ERR_INDEX);
    }
  }
  __syncthreads();
  if (s_err & PLAN_ERR_INDEX) {
    if (tid == 0) { PlanInfo o = {}; o.error = s_err; o.Cf = 0; o.max_lines_cta = 1; o.max_slots_cta = 32; o.max_items_cta = 0; o.nslots = 0; o.nitems = 0; o.has_unobserved = 0; *p.info = o; p.hdr->plan_error = s_err; }
    return;
  }
  PLAN_PHASE(0)
  // ---- B: reduced camera indices ----
  if (tid == 0) {
    int Cf = 0;
    for (int c = 0; c < MAX_CAMS; ++c) s_cam_free[c] = -1;
    for (int c = 0; c < C; ++c) {
      const bool used = (s_cam_used >> c) & 1u, cst = (s_cam_const >> c) & 1u;
      if (used && !cst) s_cam_free[c] = Cf++;
      if (!used) s_unobs = 1;
    }
    if (Cf > MAX_FREE_CAMS) s_err |= PLAN_ERR_LIMIT;
    s_Cf = Cf; s_nkeys = Cf * (Cf + 1) / 2;
  }
  __syncthreads();
  // ---- C: line_start = exclusive scan of the counts; device lines = observed lines in increasing id ----
  {
    int carry = 0, dcarry = 0;
    for (int base = 0; base < L; base += PLAN_NT) {
      const int l = base + tid;
      const int k = l < L ? p.line_cnt[l] : 0;
      if (k > 32) atomicOr(&s_err, PLAN_ERR_LIMIT);
      if (l < L && k == 0) s_unobs = 1;
      int tot, dtot;
      const int ex = plan_block_exscan(k, s_wsum, &tot);
      const int dex = plan_block_exscan(k > 0 ? 1 : 0, s_wsum, &dtot);
      if (l < L) {
        p.line_start[l] = carry + ex;
        p.fill[l] = carry + ex;
        if (k > 0) {
          p.line_gid[dcarry + dex] = l;
          s_cnt[dcarry + dex] = (unsigned char)min(k, 255);
          s_dstart[dcarry + dex] = carry + ex;
        }
      }
      carry += tot; dcarry += dtot;
    }
    if (tid == 0) { p.line_start[L] = carry; s_nd = dcarry; }
  }
  __syncthreads();
  if (s_err) {
    if (tid == 0) { PlanInfo o = {}; o.error = s_err; o.Cf = s_Cf; o.max_lines_cta = 1; o.max_slots_cta = 32; o.max_items_cta = 0; o.nslots = 0; o.nitems = 0; o.has_unobserved = s_unobs; *p.info = o; p.hdr->plan_error = s_err; }
    return;
  }
  const int nd = s_nd, Cf = s_Cf, nkeys = s_nkeys;
  PLAN_PHASE(1)
  // ---- D: group by line; inside a line ascending observation index (= the host's stable counting sort) ----
  for (int i = tid; i < N; i += PLAN_NT) p.order[atomicAdd(&p.fill[p.line_idx[i]], 1)] = i;
  __syncthreads();
  for (int l = tid; l < L; l += PLAN_NT) {
    const int b = p.line_start[l], k = p.line_cnt[l];
    if (k < 2) continue;
    int v[32];                                   // one round trip for the loads, the sort itself in thread-local memory
    for (int a = 0; a < k; ++a) v[a] = p.order[b + a];
    for (int a = 1; a < k; ++a) {
      const int x = v[a];
      int q = a;
      while (q > 0 && v[q - 1] > x) { v[q] = v[q - 1]; --q; }
      v[q] = x;
    }
    for (int a = 0; a < k; ++a) p.order[b + a] = v[a];
  }
  __syncthreads();
  PLAN_PHASE(2)
  // ---- E: lines over the CTAs of the group, balancing observation counts ----
  if (tid < CS) {
    const int r = tid;
    int li;
    if (r == CS - 1) {
      li = nd;
    } else {
      const long long target = ((long long)N * (r + 1) + CS - 1) / CS;
      int lo = 0, hi = nd;    // smallest li in [0, nd] with prefix(li) >= target; prefix(nd) = N >= target
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const long long pre = s_dstart[mid];
        if (pre >= target) hi = mid; else lo = mid + 1;
      }
      li = lo;
    }
    s_cta_line_off[r + 1] = li;
  }
  if (tid == 0) s_cta_line_off[0] = 0;
  for (int r = CS + 1 + tid; r <= MAX_G; r += PLAN_NT) s_cta_line_off[r] = nd;
  __syncthreads();
  // ---- F: whole lines into 32-lane tiles (greedy, in line order), per CTA ----
  if (tid < CS) {
    const int r = tid, lb = s_cta_line_off[r], le = s_cta_line_off[r + 1];
    int slots = 0, ln = 0;
    for (int li = lb; li < le; ++li) {
      const int k = s_cnt[li];
      if (ln + k > 32) { slots += 32 - ln; ln = 0; }
      s_line[li] = (slots & 0xffff) | (ln << 16) | (r << 22);
      slots += k; ln += k;
      if (ln == 32) ln = 0;
    }
    if (ln != 0) slots += 32 - ln;
    if (slots > 65535) atomicOr(&s_err, PLAN_ERR_LIMIT);
    atomicMax(&s_max_lines, le - lb);
    atomicMax(&s_max_slots, slots);
    s_cta_slot_off[r + 1] = slots;     // size for now, offsets after the scan below
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    s_cta_slot_off[0] = 0;
    for (int r = 0; r < CS; ++r) { const int sz = s_cta_slot_off[r + 1]; run += sz; s_cta_slot_off[r + 1] = run; }
    for (int r = CS + 1; r <= MAX_G; ++r) s_cta_slot_off[r] = run;
    s_total_slots = run;
    if (run > p.slot_cap) s_err |= PLAN_ERR_CAPACITY;
  }
  __syncthreads();
  if (s_err) {
    if (tid == 0) { PlanInfo o = {}; o.error = s_err; o.Cf = Cf; o.max_lines_cta = s_max_lines; o.max_slots_cta = s_max_slots; o.max_items_cta = 0; o.nslots = 0; o.nitems = 0; o.has_unobserved = s_unobs; *p.info = o; p.hdr->plan_error = s_err; }
    return;
  }
  const int total_slots = s_total_slots;
  PLAN_PHASE(3)
  // ---- G: slot -> device line ----
  for (int s = tid; s < total_slots; s += PLAN_NT) p.slot_line[s] = -1;
  __syncthreads();
  for (int li = tid; li < nd; li += PLAN_NT) {
    const int k = s_cnt[li], pk = s_line[li];
    const int g0 = s_cta_slot_off[pk >> 22] + (pk & 0xffff);
    for (int a = 0; a < k; ++a) p.slot_line[g0 + a] = li;
    s_mask[li] = 0u;
  }
  __syncthreads();
  PLAN_PHASE(4)
  // ---- H: slot metadata and per-line camera sets (warp = tile, lane = slot).  The observations themselves are
  // gathered into slot order by lba_gather_obs_kernel afterwards, over many CTAs: doing that copy (1.3 MB in, 1.3 MB out
  // per 10 k observations, in scattered 64-byte rows) on this one SM made this phase half of the kernel. ----
  for (int tile = warp; tile < total_slots / 32; tile += PLAN_NW) {
    const int s = tile * 32 + lane;
    const int li = p.slot_line[s];
    int cam = -1 - lane, src = -1;
    int2 m; m.x = (lane << 8) | (1 << 14); m.y = 0;
    int l = 0, k = 0, a = 0, r = 0, pk = 0;
    if (li >= 0) {
      pk = s_line[li]; k = s_cnt[li]; r = pk >> 22;
      a = s - (s_cta_slot_off[r] + (pk & 0xffff));
      src = p.order[s_dstart[li] + a];
      cam = p.cam_idx[src];
      l = p.line_idx[src];
    }
    const unsigned same = __match_any_sync(0xffffffffu, cam);
    if (li >= 0) {
      const int round = __popc(same & ((1u << lane) - 1u));
      const int lc = p.lconst[l];
      int flags = F_VALID;
      if (lc) flags |= F_LINE_FIXED;
      if ((s_cam_const >> cam) & 1u) flags |= F_CAM_FIXED;
      if (a == 0) flags |= F_HEAD;
      m.x = cam | (((pk >> 16) & 0x3f) << 8) | (k << 14) | (flags << 24);
      m.y = (li - s_cta_line_off[r]) | (round << 20);
      // the line's set of reduced cameras (constant lines: none) and where each camera's observation sits
      const int cf = lc ? -1 : s_cam_free[cam];
      if (cf >= 0) {
        const unsigned old = atomicOr(&s_mask[li], 1u << cf);
        if ((old >> cf) & 1u) atomicOr(&s_err, PLAN_DUPLICATE_CAMERA);
        s_pos[(size_t)li * pstride + cf] = (unsigned char)a;
      }
    }
    p.meta[s] = m;
    p.slot_line[s] = src;                 // from here on: slot -> caller's observation index (or -1), for the gather
  }
  __syncthreads();
  if (s_err) {
    // a camera observing the same line twice: the host planner handles that (rare; never produced by the reference)
    if (tid == 0) { PlanInfo o = {}; o.error = s_err; o.Cf = Cf; o.max_lines_cta = s_max_lines; o.max_slots_cta = s_max_slots; o.max_items_cta = 0; o.nslots = total_slots; o.nitems = 0; o.has_unobserved = s_unobs; *p.info = o; p.hdr->plan_error = s_err; }
    return;
  }
  PLAN_PHASE(5)
  // ---- I: pair blocks.  Task (r, key): the lines of CTA r seen by both cameras of the block, in line order ----
  const int kstride = nkeys + 1, ntask = CS * kstride;
  for (int t = tid; t < ntask; t += PLAN_NT) {
    const int r = t / kstride, key = t - r * kstride;
    int cnt = 0;
    if (key < nkeys) {
      int ca = 0;
      while ((ca + 1) * (ca + 2) / 2 <= key) ++ca;
      const int cb = key - ca * (ca + 1) / 2;
      if (ca != cb) {
        const unsigned need = (1u << ca) | (1u << cb);
        for (int li = s_cta_line_off[r]; li < s_cta_line_off[r + 1]; ++li) cnt += (s_mask[li] & need) == need;
      }
    }
    p.key_off[t] = cnt;
  }
  __syncthreads();
  {
    int carry = 0;
    for (int base = 0; base < ntask; base += PLAN_NT) {
      const int t = base + tid;
      const int v = t < ntask ? p.key_off[t] : 0;
      int tot;
      const int ex = plan_block_exscan(v, s_wsum, &tot);
      if (t < ntask) p.key_off[t] = carry + ex;
      carry += tot;
    }
    if (tid == 0) { s_total_items = carry; if (carry > p.item_cap) s_err |= PLAN_ERR_CAPACITY; }
  }
  __syncthreads();
  if (tid < CS) atomicMax(&s_max_items, p.key_off[tid * kstride + nkeys] - p.key_off[tid * kstride]);
  if (!s_err) {
    for (int t = tid; t < ntask; t += PLAN_NT) {
      const int r = t / kstride, key = t - r * kstride;
      if (key >= nkeys) continue;
      int ca = 0;
      while ((ca + 1) * (ca + 2) / 2 <= key) ++ca;
      const int cb = key - ca * (ca + 1) / 2;
      if (ca == cb) continue;
      const unsigned need = (1u << ca) | (1u << cb);
      int pos = p.key_off[t];
      for (int li = s_cta_line_off[r]; li < s_cta_line_off[r + 1]; ++li) {
        if ((s_mask[li] & need) != need) continue;
        const uint32_t fs = (uint32_t)(s_line[li] & 0xffff);
        const uint32_t si = fs + s_pos[(size_t)li * pstride + ca], sj = fs + s_pos[(size_t)li * pstride + cb];
        p.items[pos++] = si | (sj << 16);
      }
    }
  }
  __syncthreads();
  PLAN_PHASE(6)
  // ---- J: the header fields the solve kernel needs, and the sizes the host needs for the launch ----
  WinHdr* h = p.hdr;
  if (tid == 0) {
    h->Cf = Cf; h->n = 6 * Cf; h->nkeys = nkeys; h->vlen = lba_vlen(Cf); h->vpad = (lba_vlen(Cf) + 31) & ~31;
    h->plan_error = s_err; h->max_lines_cta = s_max_lines; h->max_slots_cta = s_max_slots; h->max_items_cta = s_max_items;
    PlanInfo o = {};
    o.error = s_err; o.Cf = Cf; o.max_lines_cta = s_max_lines; o.max_slots_cta = s_max_slots; o.max_items_cta = s_max_items;
    o.nslots = total_slots; o.nitems = s_total_items; o.has_unobserved = s_unobs;
    s_ph[7] = (int)(clock64() - t_start);
    for (int k = 0; k < 8; ++k) o.phase_cycles[k] = s_ph[k];
    *p.info = o;
  }
#undef PLAN_PHASE
  for (int r = tid; r <= MAX_G; r += PLAN_NT) { h->cta_slot_off[r] = s_cta_slot_off[r]; h->cta_line_off[r] = s_cta_line_off[r]; }
  for (int c = tid; c < MAX_CAMS; c += PLAN_NT) h->cam_free[c] = (signed char)s_cam_free[c];
}

// Observations into slot order: 4 threads per 64-byte row, rows of a window spread over gridDim.x CTAs (blockIdx.y =
// window).  Reads the slot -> observation table and the slot count the plan kernel left behind (stream-ordered after it).
__global__ void __launch_bounds__(256) lba_gather_obs_kernel(const PlanIn* __restrict__ ins) {
  const PlanIn& p = ins[blockIdx.y];
  const int total = p.info->error ? 0 : p.info->nslots * 4;
  const double2* raw = reinterpret_cast<const double2*>(p.obs_raw);
  double2* out = reinterpret_cast<double2*>(p.obs);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int src = p.slot_line[idx >> 2];
    out[idx] = src >= 0 ? raw[(size_t)src * 4 + (idx & 3)] : make_double2(0.0, 0.0);
  }
}

}  // namespace slslam
