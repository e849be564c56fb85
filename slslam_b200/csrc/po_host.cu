// Host side of the pose-graph path behind the C ABI (include/slslam_b200.h): validation, the symbolic plan
// (constant pose, reduced block indices, incident-edge CSR, the block structure of J^T J with its per-block
// contribution lists), upload, the stream-ordered LM loop (no host round trip inside), download.
// Replaces what POProblem::build + ceres::Solve do (reference src/po_problem.cpp:40-77, src/slam.cpp:1283-1293).
#include <algorithm>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <vector>

#include "common_host.h"
#include "po_kernels.cuh"

namespace slslam {

static thread_local float g_po_last_ms = 0.f;

struct PoPlan {
  int K = 0, E = 0, Kf = 0, n = 0;
  std::vector<int> slot, slot_pose, inc_off, inc, blk_i, blk_j, blk_off, contrib;
  std::vector<unsigned char> active;
  // block-sparse factorisation (symbolic part): elimination order, the blocks of L, the update list of every column
  bool sparse = false;
  int nsb = 0, max_rows = 0;
  long long dense_blocks = 0;
  std::vector<int> slot_pos, col_off, row_pos, tri_off, blk_dst, bs_chunk;
  std::vector<int2> tri;
  // level order (build_po_sparse with levels = true): columns [stage_off[s], stage_off[s + 1]) are mutually independent
  // and touch disjoint blocks of the factor, so one warp each can eliminate them side by side
  bool levels = false;
  std::vector<int> stage_off;
};

// Symbolic analysis of the block graph of the free poses: exact greedy minimum-degree order (ties by index), the
// structure of every column of L (the eliminated node's remaining neighbours, which become a clique), block numbering
// and, per column, the (destination, source a, source b) list of its updates.  Returns false when the factor is too
// full for the sparse kernel to pay (or a column exceeds its shared-memory panel): the dense path takes over.
static bool build_po_sparse(const slslam_po_desc& d, PoPlan& p, bool levels) {
  const int Kf = p.Kf;
  p.dense_blocks = (long long)Kf * (Kf + 1) / 2;
  p.levels = false; p.stage_off.clear(); p.max_rows = 0;
  if (Kf == 0) return false;
  std::vector<std::set<int> > adj(Kf);
  for (int e = 0; e < d.num_edges; ++e) {
    const int a = p.slot[d.pose_index_1[e]], b = p.slot[d.pose_index_2[e]];
    if (a >= 0 && b >= 0 && a != b) { adj[a].insert(b); adj[b].insert(a); }
  }
  std::set<std::pair<int, int> > queue;                 // (degree, node)
  for (int v = 0; v < Kf; ++v) queue.insert(std::make_pair((int)adj[v].size(), v));
  std::vector<int> order; order.reserve(Kf);
  std::vector<std::vector<int> > col_nodes(Kf);         // by elimination step: the remaining neighbours
  long long nblocks = Kf, ntri = 0;
  auto eliminate = [&](int v) -> bool {
    queue.erase(std::make_pair((int)adj[v].size(), v));
    const int step = (int)order.size();
    order.push_back(v);
    std::vector<int> nb(adj[v].begin(), adj[v].end());
    for (int u : nb) { queue.erase(std::make_pair((int)adj[u].size(), u)); adj[u].erase(v); }
    for (size_t a = 0; a < nb.size(); ++a)
      for (size_t b = a + 1; b < nb.size(); ++b) { adj[nb[a]].insert(nb[b]); adj[nb[b]].insert(nb[a]); }
    for (int u : nb) queue.insert(std::make_pair((int)adj[u].size(), u));
    adj[v].clear();
    nblocks += (long long)nb.size();
    ntri += (long long)nb.size() * ((long long)nb.size() + 1) / 2;
    p.max_rows = std::max(p.max_rows, (int)nb.size());
    col_nodes[step].swap(nb);
    if (nblocks * 3 > p.dense_blocks && Kf > 64) return false;      // more than a third of the dense factor: not sparse
    if (ntri > (1LL << 24)) return false;
    return true;
  };
  if (!levels) {
    for (int step = 0; step < Kf; ++step)
      if (!eliminate(queue.begin()->second)) return false;
  } else {
    // Multiple elimination: per level an independent set of (near-)minimum-degree nodes -- on a trajectory graph every
    // other pose of the chain -- whose eliminations do not depend on one another; inside a level, nodes that share a
    // neighbour would update the same blocks, so they are coloured apart (greedy): a (level, colour) class is a STAGE
    // whose columns the kernel eliminates concurrently, one warp each, without atomics and in a fixed order.
    p.stage_off.push_back(0);
    std::vector<int> mark(Kf, -1), colour_of(Kf, 0);
    int level = 0;
    while (!queue.empty()) {
      const int dmin = queue.begin()->first;
      const int thr = std::max(2 * dmin, 2);           // generous: the interior of a band graph has twice the degree of its ends
      std::vector<int> picked;
      for (auto it = queue.begin(); it != queue.end() && it->first <= thr; ++it) {
        const int v = it->second;
        if (mark[v] == level) continue;                  // a neighbour was picked in this level
        picked.push_back(v);
        for (int u : adj[v]) mark[u] = level;
      }
      // colours: nodes sharing a neighbour get different ones
      std::vector<std::vector<int> > used(0);
      std::map<int, std::vector<int> > nb_colours;       // neighbour -> colours taken by the picked nodes around it
      int ncol = 0;
      for (int v : picked) {
        std::vector<char> taken(ncol + 1, 0);
        for (int u : adj[v]) for (int cc : nb_colours[u]) taken[cc] = 1;
        int cc = 0;
        while (cc < ncol && taken[cc]) ++cc;
        if (cc == ncol) ++ncol;
        colour_of[v] = cc;
        for (int u : adj[v]) nb_colours[u].push_back(cc);
      }
      for (int cc = 0; cc < ncol; ++cc) {
        for (int v : picked) if (colour_of[v] == cc) { if (!eliminate(v)) return false; }
        p.stage_off.push_back((int)order.size());
      }
      ++level;
    }
    if (getenv("SLSLAM_PO_DEBUG")) {
      fprintf(stderr, "po level order: %d levels, %d stages, max rows %d, blocks %lld; columns per stage:", level, (int)p.stage_off.size() - 1, p.max_rows, nblocks);
      for (size_t k = 0; k + 1 < p.stage_off.size(); ++k) fprintf(stderr, " %d", p.stage_off[k + 1] - p.stage_off[k]);
      fprintf(stderr, "\n");
    }
    if (p.max_rows > PO_LV_MAXROWS) return false;
    // no parallelism to be had (a hub pose adjacent to everything puts every column in a stage of its own): the
    // column-at-a-time kernel is the faster one then
    if (Kf > 16 && (int)p.stage_off.size() - 1 > (3 * Kf) / 5) return false;
    p.levels = true;
  }
  if (p.max_rows > PO_SP_MAXROWS) return false;
  p.slot_pos.assign(Kf, 0);
  for (int k = 0; k < Kf; ++k) p.slot_pos[order[k]] = k;
  p.col_off.assign(Kf + 1, 0);
  p.row_pos.clear(); p.row_pos.reserve((size_t)(nblocks - Kf));
  for (int c = 0; c < Kf; ++c) {
    std::vector<int> rows;
    for (int u : col_nodes[c]) rows.push_back(p.slot_pos[u]);
    std::sort(rows.begin(), rows.end());
    p.row_pos.insert(p.row_pos.end(), rows.begin(), rows.end());
    p.col_off[c + 1] = (int)p.row_pos.size();
  }
  p.nsb = (int)nblocks;
  // block id of (row position, column position), row > column: binary search in the column's ascending row list
  auto block_id = [&](int r, int c) {
    const int* b = &p.row_pos[p.col_off[c]];
    const int* e = &p.row_pos[p.col_off[c + 1]];
    const int* it = std::lower_bound(b, e, r);
    return Kf + (int)(it - &p.row_pos[0]);
  };
  p.tri_off.assign(Kf + 1, 0);
  p.tri.clear(); p.tri.reserve((size_t)ntri);
  for (int c = 0; c < Kf; ++c) {
    const int o0 = p.col_off[c], m = p.col_off[c + 1] - o0;
    for (int a = 0; a < m; ++a)
      for (int b = 0; b <= a; ++b) {
        int2 t;
        t.x = (a == b) ? p.row_pos[o0 + a] : block_id(p.row_pos[o0 + a], p.row_pos[o0 + b]);
        t.y = a | (b << 16);
        p.tri.push_back(t);
      }
    p.tri_off[c + 1] = (int)p.tri.size();
  }
  // where the structurally non-zero blocks of J^T J go, and which of their two poses is the row block
  p.blk_dst.resize(p.blk_i.size());
  for (size_t b = 0; b < p.blk_i.size(); ++b) {
    const int si = p.blk_i[b], sj = p.blk_j[b];
    if (si == sj) { p.blk_dst[b] = p.slot_pos[si]; continue; }
    if (p.slot_pos[si] < p.slot_pos[sj]) {
      // the row block is the one eliminated later: swap the roles (and the pose1 / pose2 bit of every contribution)
      std::swap(p.blk_i[b], p.blk_j[b]);
      for (int u = p.blk_off[b]; u < p.blk_off[b + 1]; ++u) p.contrib[u] ^= 1;
    }
    p.blk_dst[b] = block_id(p.slot_pos[p.blk_i[b]], p.slot_pos[p.blk_j[b]]);
  }
  // chunks of consecutive columns (descending) whose panels fit one shared-memory stage of the back-substitution
  p.bs_chunk.clear();
  p.bs_chunk.push_back(Kf);
  for (int c = Kf; c > 0;) {
    int lo = c, blocks = 0;
    while (lo > 0 && c - lo < PO_SP_MAXROWS && blocks + (p.col_off[lo] - p.col_off[lo - 1]) <= PO_SP_MAXROWS) { blocks += p.col_off[lo] - p.col_off[lo - 1]; --lo; }
    p.bs_chunk.push_back(lo);
    c = lo;
  }
  p.sparse = true;
  return true;
}

static int validate_po(const slslam_po_desc& d) {
  if (d.num_poses < 0 || d.num_edges < 0 || d.max_iterations < 0) return SLSLAM_ERR_INVALID;
  if (d.num_edges > 0 && (!d.pose_index_1 || !d.pose_index_2 || !d.constraints)) return SLSLAM_ERR_INVALID;
  for (int e = 0; e < d.num_edges; ++e) {
    if (d.pose_index_1[e] < 0 || d.pose_index_1[e] >= d.num_poses) return SLSLAM_ERR_INVALID;
    if (d.pose_index_2[e] < 0 || d.pose_index_2[e] >= d.num_poses) return SLSLAM_ERR_INVALID;
  }
  for (size_t i = 0; i < 6 * (size_t)d.num_edges; ++i) if (!std::isfinite(d.constraints[i])) return SLSLAM_ERR_NUMERICAL;
  return SLSLAM_OK;
}

static void build_po_plan(const slslam_po_desc& d, bool all_free, PoPlan& p) {
  const int K = d.num_poses, E = d.num_edges;
  p.K = K; p.E = E;
  std::vector<char> used(K, 0);
  for (int e = 0; e < E; ++e) { used[d.pose_index_1[e]] = 1; used[d.pose_index_2[e]] = 1; }
  // pose1 of the first edge is held constant (reference src/po_problem.cpp:62-63); unused poses are never touched
  const int konst = (E > 0 && !all_free) ? d.pose_index_1[0] : -1;
  p.slot.assign(K, -1);
  p.Kf = 0;
  for (int k = 0; k < K; ++k) if (used[k] && k != konst) { p.slot[k] = p.Kf++; p.slot_pose.push_back(k); }
  p.n = 6 * p.Kf;
  p.active.assign(E, 0);
  std::vector<std::vector<int> > inc(p.Kf);
  std::map<std::pair<int, int>, std::vector<int> > blocks;
  for (int e = 0; e < E; ++e) {
    const int a = d.pose_index_1[e], b = d.pose_index_2[e];
    const int s1 = p.slot[a], s2 = p.slot[b];
    p.active[e] = (s1 >= 0 || s2 >= 0) ? 1 : 0;
    if (s1 >= 0) { inc[s1].push_back(e << 1); blocks[std::make_pair(s1, s1)].push_back(e << 1); }
    if (a != b && s2 >= 0) { inc[s2].push_back(e << 1 | 1); blocks[std::make_pair(s2, s2)].push_back(e << 1 | 1); }
    if (a != b && s1 >= 0 && s2 >= 0) {
      if (s2 > s1) blocks[std::make_pair(s2, s1)].push_back(e << 1 | 1);
      else blocks[std::make_pair(s1, s2)].push_back(e << 1);
    }
  }
  p.inc_off.assign(p.Kf + 1, 0);
  for (int s = 0; s < p.Kf; ++s) {
    p.inc_off[s + 1] = p.inc_off[s] + (int)inc[s].size();
    p.inc.insert(p.inc.end(), inc[s].begin(), inc[s].end());
  }
  p.blk_off.push_back(0);
  for (auto& kv : blocks) {
    p.blk_i.push_back(kv.first.first); p.blk_j.push_back(kv.first.second);
    p.contrib.insert(p.contrib.end(), kv.second.begin(), kv.second.end());
    p.blk_off.push_back((int)p.contrib.size());
  }
}

// One pooled device allocation; `reserve` hands out 256-byte aligned slices.
struct Pool {
  char* base = nullptr;
  size_t off = 0;
  size_t reserve(size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; }
};

// Grow-only device pool, pinned upload / result staging and events cached per host thread: a solve does no cudaMalloc,
// cudaMallocHost or cudaFree once the thread has seen a graph of that size (round 1 paid them on every call).
struct PoWorkspace {
  int device = -1;
  char* d_pool = nullptr; size_t d_cap = 0;
  char* h_up = nullptr; size_t up_cap = 0;
  char* h_res = nullptr; size_t res_cap = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> it_ev;
  int ensure(int dev, size_t d_bytes, size_t up_bytes, size_t res_bytes, int iters) {
    if (device != dev) { release(); device = dev; }
    if (!ev0) CUDA_TRY(cudaEventCreate(&ev0));
    if (!ev1) CUDA_TRY(cudaEventCreate(&ev1));
    while ((int)it_ev.size() < iters) { cudaEvent_t e; CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); it_ev.push_back(e); }
    if (d_bytes > d_cap) {
      if (d_pool) cudaFree(d_pool);
      d_pool = nullptr; d_cap = 0;
      CUDA_TRY(cudaMalloc((void**)&d_pool, d_bytes + d_bytes / 4));
      d_cap = d_bytes + d_bytes / 4;
    }
    if (up_bytes > up_cap) {
      if (h_up) cudaFreeHost(h_up);
      h_up = nullptr; up_cap = 0;
      CUDA_TRY(cudaMallocHost((void**)&h_up, up_bytes + up_bytes / 4));
      up_cap = up_bytes + up_bytes / 4;
    }
    if (res_bytes > res_cap) {
      if (h_res) cudaFreeHost(h_res);
      h_res = nullptr; res_cap = 0;
      CUDA_TRY(cudaMallocHost((void**)&h_res, res_bytes + res_bytes / 4));
      res_cap = res_bytes + res_bytes / 4;
    }
    return SLSLAM_OK;
  }
  void release() {
    if (d_pool) cudaFree(d_pool);
    if (h_up) cudaFreeHost(h_up);
    if (h_res) cudaFreeHost(h_res);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (cudaEvent_t e : it_ev) cudaEventDestroy(e);
    it_ev.clear();
    d_pool = h_up = h_res = nullptr; d_cap = up_cap = res_cap = 0; ev0 = ev1 = nullptr;
  }
};
static thread_local PoWorkspace g_po_ws;
static thread_local slslam_po_stats g_po_stats;

static int po_run(const slslam_po_desc* desc, const double* poses_in, double* poses_out, slslam_summary* summary_out,
                  double* trace_out, bool evaluate_only, double* res_out, double* j1_out, double* j2_out, double* cost_out) {
  int rc = validate_po(*desc);
  if (rc != SLSLAM_OK) return rc;
  const int K = desc->num_poses, E = desc->num_edges;
  for (size_t i = 0; i < 6 * (size_t)K; ++i) if (!std::isfinite(poses_in[i])) return SLSLAM_ERR_NUMERICAL;
  rc = ensure_device(-1);
  if (rc != SLSLAM_OK) return rc;
  int dev = 0;
  cudaGetDevice(&dev);
  PoPlan p;
  build_po_plan(*desc, evaluate_only, p);
  bool sparse = false;
  if (!evaluate_only && !getenv("SLSLAM_PO_DENSE")) {
    // level order + warp-per-column kernel when every column of that order is short enough; the minimum-degree order with
    // the column-at-a-time kernel otherwise (SLSLAM_PO_COLUMNS forces it)
    sparse = !getenv("SLSLAM_PO_COLUMNS") && build_po_sparse(*desc, p, true);
    if (!sparse) sparse = build_po_sparse(*desc, p, false);
  }
  const int n = p.n, M = n + 1, ld = (M + 7) & ~7, nblk = (int)p.blk_i.size();
  const int max_iters = desc->max_iterations;
  const int nb32 = (n + PO_NB - 1) / PO_NB;
  memset(&g_po_stats, 0, sizeof(g_po_stats));
  g_po_stats.free_poses = p.Kf; g_po_stats.sparse = sparse ? (p.levels ? 2 : 1) : 0;   // 2: level order, warp per column
  g_po_stats.factor_blocks = sparse ? p.nsb : (int64_t)p.Kf * (p.Kf + 1) / 2;
  g_po_stats.block_updates = sparse ? (int64_t)p.tri.size() : 0;
  g_po_stats.max_column_rows = sparse ? p.max_rows : p.Kf;

  Pool pool;
  const size_t Ez = std::max(E, 1), Kz = std::max(K, 1), nz = std::max(n, 1), Kfz = std::max(p.Kf, 1);
  const size_t o_idx1 = pool.reserve(4 * Ez), o_idx2 = pool.reserve(4 * Ez), o_cons = pool.reserve(48 * Ez);
  const size_t o_slot = pool.reserve(4 * Kz), o_spose = pool.reserve(4 * nz), o_act = pool.reserve(Ez);
  const size_t o_incoff = pool.reserve(4 * (p.inc_off.size() + 1)), o_inc = pool.reserve(4 * (p.inc.size() + 1));
  const size_t o_bi = pool.reserve(4 * (size_t)(nblk + 1)), o_bj = pool.reserve(4 * (size_t)(nblk + 1));
  const size_t o_boff = pool.reserve(4 * (size_t)(nblk + 2)), o_contrib = pool.reserve(4 * (p.contrib.size() + 1));
  const size_t o_spos = pool.reserve(4 * Kfz), o_coff = pool.reserve(4 * (Kfz + 1)), o_rpos = pool.reserve(4 * (p.row_pos.size() + 1));
  const size_t o_toff = pool.reserve(4 * (Kfz + 1)), o_tri = pool.reserve(8 * (p.tri.size() + 1)), o_bdst = pool.reserve(4 * (size_t)(nblk + 1));
  const size_t o_bsc = pool.reserve(4 * (p.bs_chunk.size() + 1));
  const size_t o_stg = pool.reserve(4 * (p.stage_off.size() + 1));
  const size_t o_x = pool.reserve(48 * Kz);
  const size_t o_state = pool.reserve(sizeof(PoState));
  const size_t upload_end = pool.off;
  const size_t o_xt = pool.reserve(48 * Kz);
  const size_t o_r = pool.reserve(2 * 48 * Ez), o_J1 = pool.reserve(2 * 288 * Ez), o_J2 = pool.reserve(2 * 288 * Ez);
  const size_t o_ce = pool.reserve(2 * 8 * Ez), o_mval = pool.reserve(8 * Ez);
  const size_t o_scale = pool.reserve(8 * nz), o_cn = pool.reserve(8 * nz), o_g = pool.reserve(8 * nz), o_y = pool.reserve(8 * nz);
  const size_t o_bz = pool.reserve(8 * nz), o_us = pool.reserve(8 * nz), o_yp = pool.reserve(8 * nz);
  const size_t o_summ = pool.reserve(sizeof(slslam_summary));
  const size_t o_spc = pool.reserve(64);
  const size_t o_trace = pool.reserve(8 * (size_t)SLSLAM_TRACE_WIDTH * std::max(max_iters, 1));
  const size_t o_Ld = pool.reserve(8 * (size_t)PO_NB * PO_NB * std::max(nb32, 1));
  const size_t o_flags = pool.reserve(4 * (size_t)std::max(nb32, 1));
  const size_t hb_bytes = sparse ? 288 * (size_t)std::max(p.nsb, 1) : 0;
  const size_t o_Hb = pool.reserve(hb_bytes);
  const size_t o_H = (evaluate_only || sparse) ? pool.off : pool.reserve(8 * (size_t)M * ld);
  const size_t res_bytes = 48 * Kz + sizeof(slslam_summary) + 8 * (size_t)SLSLAM_TRACE_WIDTH * std::max(max_iters, 1) + 4 * (size_t)std::max(max_iters, 1) + 1024;
  PoWorkspace& ws = g_po_ws;
  rc = ws.ensure(dev, pool.off, upload_end, res_bytes, std::max(max_iters, 1));
  if (rc != SLSLAM_OK) return rc;
  pool.base = ws.d_pool;

  char* host = ws.h_up;
  memset(host, 0, upload_end);
  if (E > 0) {
    memcpy(host + o_idx1, desc->pose_index_1, 4 * (size_t)E);
    memcpy(host + o_idx2, desc->pose_index_2, 4 * (size_t)E);
    memcpy(host + o_cons, desc->constraints, 48 * (size_t)E);
    memcpy(host + o_act, p.active.data(), (size_t)E);
  }
  if (K > 0) { memcpy(host + o_slot, p.slot.data(), 4 * (size_t)K); memcpy(host + o_x, poses_in, 48 * (size_t)K); }
  if (p.Kf > 0) memcpy(host + o_spose, p.slot_pose.data(), 4 * (size_t)p.Kf);
  memcpy(host + o_incoff, p.inc_off.data(), 4 * p.inc_off.size());
  if (!p.inc.empty()) memcpy(host + o_inc, p.inc.data(), 4 * p.inc.size());
  if (nblk > 0) {
    memcpy(host + o_bi, p.blk_i.data(), 4 * (size_t)nblk);
    memcpy(host + o_bj, p.blk_j.data(), 4 * (size_t)nblk);
    memcpy(host + o_contrib, p.contrib.data(), 4 * p.contrib.size());
  }
  memcpy(host + o_boff, p.blk_off.data(), 4 * p.blk_off.size());
  if (sparse) {
    memcpy(host + o_spos, p.slot_pos.data(), 4 * (size_t)p.Kf);
    memcpy(host + o_coff, p.col_off.data(), 4 * p.col_off.size());
    if (!p.row_pos.empty()) memcpy(host + o_rpos, p.row_pos.data(), 4 * p.row_pos.size());
    memcpy(host + o_toff, p.tri_off.data(), 4 * p.tri_off.size());
    if (!p.tri.empty()) memcpy(host + o_tri, p.tri.data(), 8 * p.tri.size());
    if (nblk > 0) memcpy(host + o_bdst, p.blk_dst.data(), 4 * (size_t)nblk);
    memcpy(host + o_bsc, p.bs_chunk.data(), 4 * p.bs_chunk.size());
    if (!p.stage_off.empty()) memcpy(host + o_stg, p.stage_off.data(), 4 * p.stage_off.size());
  }
  PoState st; memset(&st, 0, sizeof(st));
  st.radius = desc->initial_trust_region_radius > 0 ? desc->initial_trust_region_radius : 1e4;   // Ceres 1.7.0 defaults
  st.decrease_factor = 2.0;
  st.ftol = desc->function_tolerance > 0 ? desc->function_tolerance : 1e-6;
  st.gtol = desc->gradient_tolerance > 0 ? desc->gradient_tolerance : 1e-10;
  st.ptol = desc->parameter_tolerance > 0 ? desc->parameter_tolerance : 1e-8;
  st.max_iters = max_iters; st.term = SLSLAM_NO_CONVERGENCE;
  memcpy(host + o_state, &st, sizeof(st));

  PoDev d; memset(&d, 0, sizeof(d));
  char* B = pool.base;
  d.K = K; d.E = E; d.n = n; d.M = M; d.ld = ld; d.nblk = nblk;
  d.idx1 = (const int*)(B + o_idx1); d.idx2 = (const int*)(B + o_idx2); d.cons = (const double*)(B + o_cons);
  d.slot = (const int*)(B + o_slot); d.slot_pose = (const int*)(B + o_spose); d.active = (const unsigned char*)(B + o_act);
  d.inc_off = (const int*)(B + o_incoff); d.inc = (const int*)(B + o_inc);
  d.blk_i = (const int*)(B + o_bi); d.blk_j = (const int*)(B + o_bj); d.blk_off = (const int*)(B + o_boff);
  d.contrib = (const int*)(B + o_contrib);
  d.x = (double*)(B + o_x); d.xt = (double*)(B + o_xt);
  d.r = (double*)(B + o_r); d.J1 = (double*)(B + o_J1); d.J2 = (double*)(B + o_J2); d.cost_e = (double*)(B + o_ce);
  d.scale = (double*)(B + o_scale); d.cn = (double*)(B + o_cn); d.g = (double*)(B + o_g); d.y = (double*)(B + o_y);
  d.mval = (double*)(B + o_mval);
  d.H = (double*)(B + o_H); d.Ld = (double*)(B + o_Ld);
  d.st = (PoState*)(B + o_state);
  d.trace = trace_out ? (double*)(B + o_trace) : nullptr;
  d.summary = (slslam_summary*)(B + o_summ);
  d.sparse = sparse ? 1 : 0; d.Kf = p.Kf; d.nsb = p.nsb;
  d.Hb = (double*)(B + o_Hb); d.bz = (double*)(B + o_bz); d.us = (double*)(B + o_us); d.yp = (double*)(B + o_yp);
  d.slot_pos = (const int*)(B + o_spos); d.col_off = (const int*)(B + o_coff); d.row_pos = (const int*)(B + o_rpos);
  d.tri_off = (const int*)(B + o_toff); d.tri = (const int2*)(B + o_tri); d.blk_dst = (const int*)(B + o_bdst);
  d.sp_cycles = (long long*)(B + o_spc);
  d.bs_chunk = (const int*)(B + o_bsc); d.bs_nchunk = sparse ? (int)p.bs_chunk.size() - 1 : 0;
  d.stage_off = (const int*)(B + o_stg); d.nstage = (sparse && p.levels) ? (int)p.stage_off.size() - 1 : 0;

  cudaStream_t s = nullptr;
  unsigned int* d_flags = (unsigned int*)(B + o_flags);
  int bs_ctas = 1;
  // results land in pinned memory: x | summary | trace | per-iteration copies of PoState::done
  double* h_x = (double*)ws.h_res;
  slslam_summary* h_summ = (slslam_summary*)(ws.h_res + 48 * Kz);
  double* h_trace = (double*)(ws.h_res + 48 * Kz + ((sizeof(slslam_summary) + 255) & ~(size_t)255));
  volatile int* h_done = (volatile int*)((char*)h_trace + 8 * (size_t)SLSLAM_TRACE_WIDTH * std::max(max_iters, 1));
  const size_t sp_smem = PO_SP_SMEM;
  // CTAs of the level kernel's cluster: enough warps for the widest stage, at most PO_LV_CLUSTER (SLSLAM_PO_CLUSTER overrides)
  int lv_ctas = 1;
  if (sparse && p.levels) {
    int widest = 1;
    for (size_t k = 0; k + 1 < p.stage_off.size(); ++k) widest = std::max(widest, p.stage_off[k + 1] - p.stage_off[k]);
    lv_ctas = std::max(1, std::min((int)PO_LV_CLUSTER, (widest + PO_LV_NT / 32 - 1) / (PO_LV_NT / 32)));
    if (const char* ev = getenv("SLSLAM_PO_CLUSTER")) lv_ctas = std::max(1, std::min((int)PO_LV_CLUSTER, atoi(ev)));
  }
  int enqueued = 0;
  rc = SLSLAM_OK;
#define PO_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_last_error(cudaGetErrorString(e_)); cudaGetLastError(); rc = SLSLAM_ERR_CUDA; goto done; } } while (0)
  if (sparse) {
    static bool attr_set[16] = {false};
    if (dev < 0 || dev >= 16 || !attr_set[dev]) {
      PO_TRY(cudaFuncSetAttribute(po_sp_factor_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp_smem));
      PO_TRY(cudaFuncSetAttribute(po_sp_factor_levels, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PO_LV_SMEM));
      if (dev >= 0 && dev < 16) attr_set[dev] = true;
    }
  }
  PO_TRY(cudaMemcpyAsync(B, host, upload_end, cudaMemcpyHostToDevice, s));
  if (trace_out) PO_TRY(cudaMemsetAsync(B + o_trace, 0, 8 * (size_t)SLSLAM_TRACE_WIDTH * std::max(max_iters, 1), s));
  PO_TRY(cudaEventRecord(ws.ev0, s));
  if (E > 0) po_linearize<<<(PO_LIN_PARTS * E + 127) / 128, 128, 0, s>>>(d, 0, 1, evaluate_only ? 1 : 0);
  if (evaluate_only) {
    PO_TRY(cudaGetLastError());
    PO_TRY(cudaStreamSynchronize(s));
    if (E > 0) {
      PO_TRY(cudaMemcpy(res_out, d.r, 48 * (size_t)E, cudaMemcpyDeviceToHost));
      if (j1_out) PO_TRY(cudaMemcpy(j1_out, d.J1, 288 * (size_t)E, cudaMemcpyDeviceToHost));
      if (j2_out) PO_TRY(cudaMemcpy(j2_out, d.J2, 288 * (size_t)E, cudaMemcpyDeviceToHost));
    }
    if (cost_out) {
      std::vector<double> ce(Ez, 0.0);
      if (E > 0) PO_TRY(cudaMemcpy(ce.data(), d.cost_e, 8 * (size_t)E, cudaMemcpyDeviceToHost));
      double c = 0.0;
      for (int e = 0; e < E; ++e) c += ce[e];
      *cost_out = c;
    }
    goto done;
  }
  if (!sparse && n > 0) {
    // dense path: back-substitution spreads its 32-column blocks over co-resident CTAs (cooperative launch)
    int sms = 1, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, po_backsolve, 256, 0) != cudaSuccess) { cudaGetLastError(); per_sm = 1; }
    bs_ctas = std::max(1, std::min(nb32, sms * std::max(per_sm, 1)));
    bs_ctas = std::min(bs_ctas, 64);
    if ((nb32 + bs_ctas - 1) / bs_ctas > PO_BS_MAXOWN) {
      set_last_error("pose graph with a nearly full factor and more than 2730 free poses: beyond the dense fallback (slslam_po_get_limits)");
      rc = SLSLAM_ERR_UNSUPPORTED; goto done;
    }
    PO_TRY(cudaMemsetAsync(d_flags, 0, 4 * (size_t)std::max(nb32, 1), s));
  }
  if (n > 0) po_colnorm_grad<<<(n + 127) / 128, 128, 0, s>>>(d, 0);
  po_refresh<<<1, 256, 0, s>>>(d, 1);
  if (n > 0) {
    for (int it = 0; it < max_iters; ++it) {
      // PoState::done of every iteration is copied to pinned memory behind it.
      if (it >= 3) {
        // the host stays at most three iterations ahead of the device (their launches are already queued, so the device
        // never waits for the host) and stops enqueueing once an iteration it has seen complete reports termination
        PO_TRY(cudaEventSynchronize(ws.it_ev[it - 3]));
        if (h_done[it - 3]) break;
      }
      if (sparse) {
        PO_TRY(cudaMemsetAsync(d.Hb, 0, hb_bytes, s));
        po_sp_assemble<<<(nblk * 36 + n + 255) / 256, 256, 0, s>>>(d);
        if (p.levels) {
          // one thread-block cluster: the columns of a stage spread over the warps of up to 8 SMs of one GPC
          cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
          cfg.gridDim = dim3((unsigned)lv_ctas); cfg.blockDim = dim3(PO_LV_NT); cfg.dynamicSmemBytes = PO_LV_SMEM; cfg.stream = s;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = (unsigned)lv_ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          cudaError_t le = cudaLaunchKernelEx(&cfg, po_sp_factor_levels, d);
          if (le != cudaSuccess && lv_ctas > 1) {
            // no room for the cluster right now (or clusters unavailable): the same kernel as a single CTA
            cudaGetLastError();
            lv_ctas = 1;
            cfg.gridDim = dim3(1); at[0].val.clusterDim.x = 1;
            le = cudaLaunchKernelEx(&cfg, po_sp_factor_levels, d);
          }
          PO_TRY(le);
        } else {
          po_sp_factor_solve<<<1, PO_SP_NT, sp_smem, s>>>(d);
        }
      } else {
        PO_TRY(cudaMemsetAsync(d.H, 0, 8 * (size_t)M * ld, s));
        po_assemble<<<(nblk * 36 + n + 255) / 256, 256, 0, s>>>(d);
        for (int k0 = 0; k0 < n; k0 += PO_NB) {
          const int nb = std::min(PO_NB, n - k0), t0 = k0 + nb;
          po_chol_panel<<<1 + (M - t0 + PO_TR - 1) / PO_TR, PO_TR, 0, s>>>(d, k0);
          if (t0 < n) {
            const int T = (M - t0 + PO_TS - 1) / PO_TS;
            po_chol_syrk<<<T * (T + 1) / 2, 256, 0, s>>>(d, k0, nb, t0, T);
          }
        }
        unsigned int gen = (unsigned int)(it + 1);
        void* args[3] = {(void*)&d, (void*)&d_flags, (void*)&gen};
        PO_TRY(cudaLaunchCooperativeKernel((const void*)po_backsolve, dim3((unsigned)bs_ctas), dim3(256), args, 0, s));
      }
      po_step<<<(6 * K + E + 255) / 256, 256, 0, s>>>(d);
      po_linearize<<<(PO_LIN_PARTS * E + 127) / 128, 128, 0, s>>>(d, 1, 0, 0);
      po_decide<<<1, 256, 0, s>>>(d);
      po_accept<<<(6 * K + 255) / 256, 256, 0, s>>>(d);
      po_colnorm_grad<<<(n + 127) / 128, 128, 0, s>>>(d, 1);
      po_refresh<<<1, 256, 0, s>>>(d, 0);
      PO_TRY(cudaMemcpyAsync((void*)(h_done + it), &d.st->done, 4, cudaMemcpyDeviceToHost, s));
      PO_TRY(cudaEventRecord(ws.it_ev[it], s));
      ++enqueued;
    }
  }
  g_po_stats.iterations_enqueued = enqueued;
  po_finish<<<1, 1, 0, s>>>(d);
  PO_TRY(cudaGetLastError());
  PO_TRY(cudaEventRecord(ws.ev1, s));
  {
    slslam_summary summ;
    PO_TRY(cudaMemcpyAsync(h_x, d.x, 48 * (size_t)Kz, cudaMemcpyDeviceToHost, s));
    PO_TRY(cudaMemcpyAsync(h_summ, d.summary, sizeof(summ), cudaMemcpyDeviceToHost, s));
    if (trace_out) PO_TRY(cudaMemcpyAsync(h_trace, d.trace, 8 * (size_t)SLSLAM_TRACE_WIDTH * std::max(max_iters, 0), cudaMemcpyDeviceToHost, s));
    PO_TRY(cudaStreamSynchronize(s));
    PO_TRY(cudaEventElapsedTime(&g_po_last_ms, ws.ev0, ws.ev1));
    // parameters are only overwritten once everything has succeeded
    if (sparse) { long long cyc[4] = {0, 0, 0, 0}; if (cudaMemcpy(cyc, d.sp_cycles, 32, cudaMemcpyDeviceToHost) == cudaSuccess) for (int k = 0; k < 4; ++k) g_po_stats.factor_cycles[k] = cyc[k]; }
    if (K > 0) memcpy(poses_out, h_x, 48 * (size_t)K);
    if (trace_out) memcpy(trace_out, h_trace, 8 * (size_t)SLSLAM_TRACE_WIDTH * std::max(max_iters, 0));
    if (summary_out) *summary_out = *h_summ;
  }
done:
#undef PO_TRY
  return rc;
}

}  // namespace slslam

using namespace slslam;

extern "C" {

int slslam_po_solve_trace(const slslam_po_desc* desc, double* poses_inout, slslam_summary* summary_out, double* trace_out) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!desc || (!poses_inout && desc->num_poses > 0)) return SLSLAM_ERR_INVALID;
  return po_run(desc, poses_inout, poses_inout, summary_out, trace_out, false, nullptr, nullptr, nullptr, nullptr);
}

int slslam_po_solve(const slslam_po_desc* desc, double* poses_inout, slslam_summary* summary_out) {
  return slslam_po_solve_trace(desc, poses_inout, summary_out, nullptr);
}

int slslam_po_evaluate(const slslam_po_desc* desc, const double* poses, double* residuals, double* jac_pose1,
                       double* jac_pose2, double* cost_out) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!desc || !poses || !residuals) return SLSLAM_ERR_INVALID;
  return po_run(desc, poses, nullptr, nullptr, nullptr, true, residuals, jac_pose1, jac_pose2, cost_out);
}

float slslam_po_last_solve_ms(void) { return g_po_last_ms; }

void slslam_po_last_stats(slslam_po_stats* out) { if (out) *out = g_po_stats; }

int slslam_po_plan_check(const slslam_po_desc* desc, int32_t force_columns, slslam_po_plan_info* out) {
  set_last_error("");
  if (!desc || !out) return SLSLAM_ERR_INVALID;
  int rc = validate_po(*desc);
  if (rc != SLSLAM_OK) return rc;
  PoPlan p;
  build_po_plan(*desc, false, p);
  bool sparse = !force_columns && build_po_sparse(*desc, p, true);
  if (!sparse) sparse = build_po_sparse(*desc, p, false);
  memset(out, 0, sizeof(*out));
  out->free_poses = p.Kf;
  if (!sparse) return SLSLAM_OK;
  out->order = p.levels ? 2 : 1;
  out->factor_blocks = p.nsb; out->block_updates = (int64_t)p.tri.size(); out->max_column_rows = p.max_rows;
  const int Kf = p.Kf;
  auto fail = [&](const char* what) { set_last_error(what); return SLSLAM_ERR_NUMERICAL; };
  // structure: rows of a column come later in the order, ascending; positions are a permutation
  std::vector<char> seen(Kf, 0);
  for (int s = 0; s < Kf; ++s) { const int c = p.slot_pos[s]; if (c < 0 || c >= Kf || seen[c]) return fail("positions are not a permutation"); seen[c] = 1; }
  for (int c = 0; c < Kf; ++c)
    for (int k = p.col_off[c]; k < p.col_off[c + 1]; ++k) {
      if (p.row_pos[k] <= c || p.row_pos[k] >= Kf) return fail("a row of a column is not eliminated after it");
      if (k > p.col_off[c] && p.row_pos[k] <= p.row_pos[k - 1]) return fail("rows of a column are not ascending");
    }
  // every update targets the diagonal block of its row or an existing off-diagonal block (row a of column row b)
  for (int c = 0; c < Kf; ++c) {
    const int o0 = p.col_off[c], m = p.col_off[c + 1] - o0;
    if (p.tri_off[c + 1] - p.tri_off[c] != m * (m + 1) / 2) return fail("update list of a column has the wrong length");
    for (int t = p.tri_off[c]; t < p.tri_off[c + 1]; ++t) {
      const int a = p.tri[t].y & 0xffff, b = p.tri[t].y >> 16, dst = p.tri[t].x;
      if (a >= m || b > a) return fail("update sources out of range");
      const int ra = p.row_pos[o0 + a], rb = p.row_pos[o0 + b];
      if (a == b) { if (dst != ra) return fail("diagonal update does not target its row's pivot block"); continue; }
      if (dst < Kf || dst >= p.nsb) return fail("off-diagonal update outside the block range");
      const int k = dst - Kf;
      if (k < p.col_off[rb] || k >= p.col_off[rb + 1] || p.row_pos[k] != ra) return fail("off-diagonal update targets the wrong block");
    }
  }
  // every block of J^T J has a home
  for (size_t b = 0; b < p.blk_dst.size(); ++b) if (p.blk_dst[b] < 0 || p.blk_dst[b] >= p.nsb) return fail("a block of J^T J has no home in the factor");
  if (p.levels) {
    out->stages = (int)p.stage_off.size() - 1;
    if (p.stage_off.empty() || p.stage_off.front() != 0 || p.stage_off.back() != Kf) return fail("stages do not cover the columns");
    std::vector<int> owner_blk(p.nsb, -1), owner_row(Kf, -1);
    for (int s = 0; s + 1 < (int)p.stage_off.size(); ++s) {
      const int c0 = p.stage_off[s], c1 = p.stage_off[s + 1];
      if (c1 <= c0) return fail("empty stage");
      out->widest_stage = std::max(out->widest_stage, c1 - c0);
      for (int c = c0; c < c1; ++c) {
        // no column of the stage is a row of another one (independence), destinations and rhs rows are private
        for (int k = p.col_off[c]; k < p.col_off[c + 1]; ++k) {
          if (p.row_pos[k] >= c0 && p.row_pos[k] < c1) return fail("two columns of a stage are adjacent");
          if (owner_row[p.row_pos[k]] == s) return fail("two columns of a stage update the same right-hand-side rows");
        }
        for (int k = p.col_off[c]; k < p.col_off[c + 1]; ++k) owner_row[p.row_pos[k]] = s;
        for (int t = p.tri_off[c]; t < p.tri_off[c + 1]; ++t) {
          int& o = owner_blk[p.tri[t].x];
          if (o == s * Kf + c) continue;
          if (o >= s * Kf && o < (s + 1) * Kf) return fail("two columns of a stage update the same block");
          o = s * Kf + c;
        }
      }
    }
  }
  return SLSLAM_OK;
}

void slslam_po_get_limits(slslam_po_limits* out) {
  if (!out) return;
  out->max_column_blocks_sparse = PO_SP_MAXROWS;
  out->max_free_poses_dense = PO_NB * PO_BS_MAXOWN * 64 / 6;
}

}  // extern "C"
