// optgraph.cpp — cache-trace analysis of the Graph module, plus its device-free C ABI.
//
// A cache trace starts at a non-leaf node whose value is consumed by at most one later
// node of the same length, and follows that single consumer forward for as long as the
// chain keeps that property (the first node that breaks it still joins the trace).
// Buffers on one trace never need to be alive at the same time, so they may share one
// allocation (optimize_mem_graph) and, when every op on it is unary, collapse into one
// kernel (unary_fusing).  Restates src/modules/graph/opt_graph/optimize.rs:19-132.
#include "optgraph.h"

#include <algorithm>
#include <memory>

#include "common.h"

namespace cb {

bool GraphNode::is_leaf() const
{
    return std::all_of(deps.begin(), deps.end(), [this](size_t d) { return d == idx; });
}

size_t OptGraph::add_leaf(size_t len) { return add_node(len, {}); }

size_t OptGraph::add_node(size_t len, std::vector<size_t> deps)
{
    GraphNode n;
    n.idx = nodes_.size();
    n.deps = std::move(deps);
    n.len = len;
    nodes_.push_back(std::move(n));
    return nodes_.back().idx;
}

static bool depends_on(const GraphNode &n, size_t idx)
{
    return std::find(n.deps.begin(), n.deps.end(), idx) != n.deps.end();
}

bool OptGraph::is_path_optimizable(size_t idx) const
{
    const GraphNode &at = nodes_[idx];
    if (at.is_leaf()) return false;
    int consumers = 0;
    for (size_t k = idx + 1; k < nodes_.size(); k++) {
        const GraphNode &c = nodes_[k];
        if (c.len != at.len || !depends_on(c, idx)) continue;
        if (++consumers > 1) return false;
    }
    return true;
}

std::vector<size_t> OptGraph::trace_cache_path_raw(size_t start) const
{
    std::vector<size_t> trace;
    if (!is_path_optimizable(start)) return trace;
    size_t cur = start;
    for (size_t k = start + 1; k < nodes_.size(); k++) {
        const GraphNode &c = nodes_[k];
        if (c.skip || !depends_on(c, cur) || c.len != nodes_[start].len) continue;
        cur = k;
        trace.push_back(cur);
        if (!is_path_optimizable(cur)) break;
    }
    return trace;
}

std::vector<CacheTrace> OptGraph::cache_traces() const
{
    std::vector<CacheTrace> traces;
    std::vector<char> visited(nodes_.size(), 0);
    for (const GraphNode &n : nodes_) {
        if (n.is_leaf() || n.skip || visited[n.idx]) continue;
        const std::vector<size_t> raw = trace_cache_path_raw(n.idx);
        if (raw.empty()) continue;
        CacheTrace t;
        t.cache_idx = n.idx;
        for (size_t i : raw) {
            if (visited[i]) continue;
            visited[i] = 1;
            t.use_cache_idxs.push_back(i);
        }
        traces.push_back(std::move(t));
    }
    return traces;
}

}  // namespace cb

// ------------------------------------------------------------------ C ABI
struct cb_optgraph {
    cb::OptGraph g;
};

using cb::fail;

extern "C" int32_t cb_optgraph_create(cb_optgraph **out)
{
    CB_CHECK_ARG(out, "out is null");
    *out = new cb_optgraph();
    return CB_OK;
}
extern "C" int32_t cb_optgraph_destroy(cb_optgraph *g)
{
    delete g;
    return CB_OK;
}
extern "C" int32_t cb_optgraph_add_leaf(cb_optgraph *g, size_t len, int64_t *idx)
{
    CB_CHECK_ARG(g, "null graph");
    const size_t i = g->g.add_leaf(len);
    if (idx) *idx = (int64_t)i;
    return CB_OK;
}
extern "C" int32_t cb_optgraph_add_node(cb_optgraph *g, size_t len, const int64_t *deps, int32_t n_deps, int64_t *idx)
{
    CB_CHECK_ARG(g && (n_deps == 0 || deps) && n_deps >= 0, "bad argument");
    std::vector<size_t> d;
    for (int32_t i = 0; i < n_deps; i++) {
        if (deps[i] < 0 || (size_t)deps[i] > g->g.size()) return fail(CB_ERR_INVALID_ARG, "dependency %lld does not exist", (long long)deps[i]);
        d.push_back((size_t)deps[i]);
    }
    const size_t i = g->g.add_node(len, std::move(d));
    if (idx) *idx = (int64_t)i;
    return CB_OK;
}
static int32_t check_idx(cb_optgraph *g, int64_t idx)
{
    if (!g) return fail(CB_ERR_INVALID_ARG, "null graph");
    if (idx < 0 || (size_t)idx >= g->g.size()) return fail(CB_ERR_INVALID_ARG, "node %lld does not exist", (long long)idx);
    return CB_OK;
}
extern "C" int32_t cb_optgraph_set_skip(cb_optgraph *g, int64_t idx, int32_t skip)
{
    CB_TRY(check_idx(g, idx));
    g->g.node((size_t)idx).skip = skip != 0;
    return CB_OK;
}
extern "C" int32_t cb_optgraph_is_path_optimizable(cb_optgraph *g, int64_t idx, int32_t *out)
{
    CB_TRY(check_idx(g, idx));
    CB_CHECK_ARG(out, "out is null");
    *out = g->g.is_path_optimizable((size_t)idx) ? 1 : 0;
    return CB_OK;
}
extern "C" int32_t cb_optgraph_trace_cache_path_raw(cb_optgraph *g, int64_t idx, int64_t *out, size_t cap, size_t *written)
{
    CB_TRY(check_idx(g, idx));
    const std::vector<size_t> t = g->g.trace_cache_path_raw((size_t)idx);
    if (written) *written = t.size();
    if (t.size() > cap) return fail(CB_ERR_INVALID_ARG, "output too small (%zu needed)", t.size());
    for (size_t i = 0; i < t.size(); i++) out[i] = (int64_t)t[i];
    return CB_OK;
}

int32_t cb_flatten_traces(const std::vector<cb::CacheTrace> &traces, int64_t *out, size_t cap, size_t *written)
{
    size_t need = 0;
    for (const auto &t : traces) need += 2 + t.use_cache_idxs.size();
    if (written) *written = need;
    if (need > cap) return fail(CB_ERR_INVALID_ARG, "output too small (%zu needed)", need);
    size_t w = 0;
    for (const auto &t : traces) {
        out[w++] = (int64_t)t.cache_idx;
        out[w++] = (int64_t)t.use_cache_idxs.size();
        for (size_t i : t.use_cache_idxs) out[w++] = (int64_t)i;
    }
    return CB_OK;
}

extern "C" int32_t cb_optgraph_cache_traces(cb_optgraph *g, int64_t *out, size_t cap, size_t *written)
{
    CB_CHECK_ARG(g, "null graph");
    return cb_flatten_traces(g->g.cache_traces(), out, cap, written);
}
