// optgraph.h — the dependency graph behind the Graph module and its cache-trace analysis
// (src/modules/graph/node.rs:2-42, opt_graph.rs:6-41, opt_graph/optimize.rs:19-132).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace cb {

struct GraphNode {
    size_t idx = 0;
    std::vector<size_t> deps;
    size_t len = 0;
    bool skip = false;  // set by Buffer::checkpoint(): never aliased or fused through

    // a node without dependencies, or one that only depends on itself (node.rs:36-42)
    bool is_leaf() const;
};

// which buffers of the graph may share one allocation / be fused into one kernel
struct CacheTrace {
    size_t cache_idx = 0;
    std::vector<size_t> use_cache_idxs;
};

class OptGraph {
public:
    size_t add_leaf(size_t len);
    size_t add_node(size_t len, std::vector<size_t> deps);
    const GraphNode &node(size_t idx) const { return nodes_[idx]; }
    GraphNode &node(size_t idx) { return nodes_[idx]; }
    size_t size() const { return nodes_.size(); }

    bool is_path_optimizable(size_t idx) const;
    std::vector<size_t> trace_cache_path_raw(size_t idx) const;
    std::vector<CacheTrace> cache_traces() const;

private:
    std::vector<GraphNode> nodes_;
};

}  // namespace cb
