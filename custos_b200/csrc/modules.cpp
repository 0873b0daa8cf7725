// modules.cpp — custos' module stack (Base, Cached, Lazy, Graph, Autograd) restated in C++
// around the CUDA device, exported through the cbm_* part of the C ABI.
//
// In the reference the stack is a compile-time nest of mixins, `CUDA<Graph<Lazy<Base>>>`;
// here it is one device object configured by a bit mask, because the composition has to be
// driven through a C ABI.  The observable behaviour follows the reference module by module:
//
//   Base      eager: add_op runs the operation at once, retrieve allocates
//             (src/modules/base.rs:53-62,118-129)
//   Cached    the k-th retrieve after a cursor reset returns the k-th allocation
//             (src/modules/cached.rs:173-229, src/range.rs:35-48)
//   Lazy      retrieve hands out ids and defers the allocation, add_op records, run()
//             allocates and replays (src/modules/lazy.rs:93-107,143-198,325-387)
//   Graph     every retrieve adds a node (len, parent nodes); cache traces drive buffer
//             aliasing and unary fusing (src/modules/graph.rs:79-127,159-195,
//             src/modules/lazy/optimization.rs:4-95, src/devices/fusing.rs:41-92)
//   Autograd  requires-grad propagation, a tape of grad functions replayed in reverse,
//             gradient buffers allocated (zeroed) on first use (src/modules/autograd.rs:108-281,
//             autograd/tape.rs:30-83, autograd/gradients.rs:110-124, buffer/impl_autograd.rs:20-56)
//
// Deliberate differences from the reference (each keeps the CPU-visible results):
//   * run() does not launch every op eagerly AND replay a captured graph;
//     with graph replay enabled the ops are captured once and only replayed.
//   * backward() seeds the output gradient with a device-side fill instead of a host vector
//     of ones (impl_autograd.rs:32).
//   * optimize_mem_graph keeps the deferred allocation of buffers that are on no cache trace
//     (the reference drains and forgets them, lazy/optimization.rs:11, and run() then fails).
//   * unary_fusing fuses maximal runs of unary-hinted ops on a trace and leaves other ops of
//     the trace alone (the reference turns a trace that contains a non-unary op into no-ops,
//     lazy/optimization.rs:77-91); on pure unary traces — all the reference tests — both agree.
//   * aliasing / fusing skip buffers whose dtype or length differ instead of panicking later.
//   * Autograd + fusing / aliasing.  In the reference, `unary_ew` records grad functions that read the INPUT
//     buffer of every op (src/unary.rs:118-128) while unary fusing turns the ops writing those buffers into no-ops
//     (lazy/optimization.rs:77-91) and optimize_mem_graph lets them share one allocation (lazy/optimization.rs:4-44):
//     backward() then reads zeros / overwritten values and returns wrong gradients without an error.  Here both
//     passes look at the tape first (tape_plan_for_chain): when the K grad functions of a chain are on the tape they
//     are replaced by ONE chain-grad kernel that recomputes every intermediate from the chain's input
//     (CB_KERNEL_CHAIN_GRAD: same roundings in the same order as the K add_unary_grad calls, with the
//     intermediate gradients — which no longer have a buffer worth the name — taken as zero-initialised); when
//     only part of the chain is on the tape the pass leaves that chain alone.  Forward + backward of a fused
//     chain is then 2 kernels whatever K is.
#include <algorithm>
#include <cstring>
#include <functional>
#include <memory>
#include <unordered_map>
#include <unordered_set>

#include "device.h"
#include "expr.h"
#include "optgraph.h"

using namespace cb;

int32_t cb_flatten_traces(const std::vector<cb::CacheTrace> &traces, int64_t *out, size_t cap, size_t *written);

namespace {

// what Lazy::buffers / Autograd::no_grads_pool hold: id -> shallow view of the storage
struct Entry {
    uint64_t ptr = 0;
    size_t len = 0;
    int32_t dtype = CB_F32;
};

// a user visible Buffer
struct Handle {
    uint64_t id = 0;   // HasId::id(): device address for eager buffers (cuda_ptr.rs:42-50), cursor for lazy ones
    size_t len = 0;
    int32_t dtype = CB_F32;
    uint64_t ptr = 0;  // storage owned by (or lent to) this handle; 0 for lazily retrieved buffers
    bool lazy = false;
    bool owned = false;  // freed on drop (AllocFlag::None); cached and gradient buffers are not
    // sharded device: `len` is this rank's slice [shard_begin, shard_begin + len) of a buffer of global_len elements
    size_t global_len = 0, shard_begin = 0;
};

enum class OpKind { NoOp, Apply, UnaryGrad, Binary, Apply2, Clear };

// Operation (src/modules/lazy/lazy_graph.rs:8-12): argument ids + what to launch + the op hint
struct Op {
    OpKind kind = OpKind::NoOp;
    std::vector<uint64_t> arg_ids;
    int32_t dtype = CB_F32;
    cb_expr *expr = nullptr;
    int32_t binop = 0;
    bool unary_hint = false;       // OpHint::Unary (src/op_hint.rs:5-11)
    bool fused = false;            // OpHint::UnaryFused
    std::vector<cb_node> hint;     // the closure of the hint, as IR
    // what the op computes as ONE expression tree over its input buffers (X = in[0], Y = in[1]);
    // kept for Apply and Binary ops so that element-wise fusing can splice producers into consumers
    std::vector<uint64_t> in;
    uint64_t out = 0;
    std::vector<cb_node> tree;
};

cb_node mk_node(int32_t op, int32_t a = -1, int32_t b = -1)
{
    cb_node n;
    std::memset(&n, 0, sizeof n);
    n.op = op;
    n.a = a;
    n.b = b;
    return n;
}

// `outer` with every `which` marker (CB_OP_X / CB_OP_Y) replaced by the value of `inner`; inner's markers
// become inner_map[0/1], outer's other marker becomes outer_other_to.
std::vector<cb_node> substitute(const std::vector<cb_node> &outer, int32_t which, int32_t outer_other_to,
                                const std::vector<cb_node> &inner, const int32_t inner_map[2])
{
    std::vector<cb_node> r;
    r.reserve(inner.size() + outer.size());
    for (const cb_node &n : inner) {
        cb_node c = n;
        if (c.op == CB_OP_X) c.op = inner_map[0];
        else if (c.op == CB_OP_Y) c.op = inner_map[1];
        r.push_back(c);
    }
    const int32_t inner_root = (int32_t)r.size() - 1;
    std::vector<int32_t> idx(outer.size(), -1);
    for (size_t i = 0; i < outer.size(); i++) {
        cb_node c = outer[i];
        if (c.op == which) {
            idx[i] = inner_root;
            continue;
        }
        if (c.op == CB_OP_X || c.op == CB_OP_Y) c.op = outer_other_to;
        if (c.a >= 0) c.a = idx[(size_t)c.a];
        if (c.b >= 0) c.b = idx[(size_t)c.b];
        idx[i] = (int32_t)r.size();
        r.push_back(c);
    }
    if (idx.back() != (int32_t)r.size() - 1) r.push_back(mk_node(CB_OP_IDENTITY, idx.back()));  // outer was just the marker
    return r;
}

// a grad function on the tape: the closure of unary_ew (src/unary.rs:118-128)
struct GradOp {
    uint64_t buf_id = 0, out_id = 0;
    int32_t dtype = CB_F32;
    cb_expr *grad_expr = nullptr;            // CB_KERNEL_UNARY_GRAD, or CB_KERNEL_CHAIN_GRAD when `chain`
    std::vector<cb_node> fwd_ir, grad_ir;    // the two closures of unary_ew, kept so that a chain can be fused later
    bool chain = false;                      // replaces the K grad functions of a fused chain buf_id -> ... -> out_id
    size_t chain_len = 1;
};

struct Deferred {
    uint64_t id;
    size_t len;
    int32_t dtype;
};

struct CacheSlot {
    uint64_t ptr;
    size_t len;
    int32_t dtype;
};

}  // namespace

struct cbm_device {
    cb_device *raw = nullptr;
    cb_comm *comm = nullptr;  // set by cbm_device_set_comm: buffers are slices, reductions combine over the ranks
    uint32_t mods = 0;
    int32_t dtype = CB_F32;  // the module's `T` (Lazy<Mods, T = f32>, Graph<Mods, T = f32>)

    uint64_t next_handle = 1;
    std::unordered_map<uint64_t, Handle> handles;
    std::unordered_map<uint64_t, Entry> buffers;  // id -> storage
    std::vector<uint64_t> lazy_allocs;            // memory owned by the Lazy module

    // Cursor (features.rs:68-111): shared by Cached / Lazy, read by Graph
    uint64_t cursor = 0;
    std::unordered_map<uint64_t, CacheSlot> cache;  // Cached: cursor -> allocation
    std::vector<uint64_t> cache_orphans;            // allocations replaced by aliasing, freed with the device

    // Lazy
    bool lazy_enabled = true;
    std::vector<Deferred> alloc_later;
    std::unordered_set<uint64_t> allocated_ids;
    std::vector<Op> ops;
    std::unordered_map<uint64_t, size_t> op_of_id;  // retrieved buffer id -> index of the op writing it
    bool replay_enabled = false;
    cb_graph *replay = nullptr;
    bool replay_valid = false;

    // Graph (graph_translator.rs:9-18)
    OptGraph graph;
    std::unordered_map<uint64_t, size_t> buf_id_to_idx;
    std::unordered_map<size_t, uint64_t> idx_to_buf_id;
    std::unordered_map<size_t, uint64_t> idx_to_cursor;
    std::unordered_set<uint64_t> contains_ids;

    // Autograd
    bool grad_enabled = true;
    std::vector<GradOp> tape;
    std::unordered_map<uint64_t, Entry> grads;          // Gradients::grads_pool
    std::unordered_map<uint64_t, uint64_t> grad_handle; // id -> handle lent to the user
    std::unordered_map<uint64_t, bool> requires_grad;   // Gradients::buf_requires_grad

    bool has(uint32_t m) const { return (mods & m) != 0; }
    bool recording() const { return has(CBM_LAZY) && lazy_enabled; }
    void invalidate_replay() { replay_valid = false; }

    Handle *handle(cbm_buf b)
    {
        auto it = handles.find(b);
        return it == handles.end() ? nullptr : &it->second;
    }
    // Buffer::replace(): the storage currently registered under the buffer's id
    const Entry *resolve(uint64_t id) const
    {
        auto it = buffers.find(id);
        return it == buffers.end() ? nullptr : &it->second;
    }
    std::unordered_map<uint64_t, int> id_refs;  // live handles per id (cached slots are handed out repeatedly)
    cbm_buf new_handle(const Handle &h)
    {
        const cbm_buf b = next_handle++;
        handles[b] = h;
        id_refs[h.id]++;
        return b;
    }
};

namespace {

#define GET_HANDLE(var, d, b)                                                                      \
    Handle *var = (d)->handle(b);                                                                  \
    if (!var) return fail(CB_ERR_INVALID_ARG, "%s: unknown buffer handle %llu", __func__, (unsigned long long)(b))

int32_t alloc_zeroed(cbm_device *d, int32_t dtype, size_t len, uint64_t *ptr)
{
    if (len == 0) return fail(CB_ERR_ZERO_LENGTH, "zero length buffer");  // DeviceError::ZeroLengthBuffer
    return cb_alloc(d->raw, len * dtype_size(dtype), 1, ptr);
}

// Graph::on_new_buffer (graph.rs:115-127) + Lazy::on_new_buffer (lazy.rs:200-214)
void on_new_buffer(cbm_device *d, const Handle &h)
{
    if (d->has(CBM_GRAPH)) {
        d->buf_id_to_idx[h.id] = d->graph.size();
        d->graph.add_leaf(h.len);
    }
    d->buffers[h.id] = Entry{h.ptr, h.len, h.dtype};
    if (d->has(CBM_AUTOGRAD)) d->requires_grad[h.id] = false;  // `insert`, not `entry().or_insert` (autograd.rs:86-91)
}

// Lazy::alloc_later (lazy.rs:175-181 + the callback at :345-380)
int32_t run_alloc_later(cbm_device *d)
{
    std::vector<Deferred> todo;
    todo.swap(d->alloc_later);
    for (const Deferred &a : todo) {
        if (d->allocated_ids.count(a.id)) continue;
        if (d->buffers.count(a.id)) return fail(CB_ERR_STATE, "IDs collided! Maybe pointing address already occupied this ID.");
        uint64_t p = 0;
        CB_TRY(alloc_zeroed(d, a.dtype, a.len, &p));
        d->lazy_allocs.push_back(p);
        d->allocated_ids.insert(a.id);
        d->buffers[a.id] = Entry{p, a.len, a.dtype};
    }
    return CB_OK;
}

// runs one recorded operation now; every argument is looked up by id (any_op.rs:112-124)
int32_t call_op(cbm_device *d, const Op &op)
{
    if (op.kind == OpKind::NoOp) return CB_OK;
    const Entry *arg[3] = {nullptr, nullptr, nullptr};
    for (size_t i = 0; i < op.arg_ids.size() && i < 3; i++) {
        arg[i] = d->resolve(op.arg_ids[i]);
        if (!arg[i]) return fail(CB_ERR_INVALID_LAZY_BUF, "InvalidLazyBuf: buffer id %llu is not alive",
                                 (unsigned long long)op.arg_ids[i]);
    }
    switch (op.kind) {
    case OpKind::Apply:  // args: (out, in)
        return cb_apply(d->raw, op.expr, arg[1]->ptr, arg[0]->ptr, std::min(arg[0]->len, arg[1]->len));
    case OpKind::UnaryGrad:  // args: (lhs, lhs_grad, out)
        return cb_unary_grad(d->raw, op.expr, arg[0]->ptr, arg[1]->ptr, arg[2]->ptr, arg[0]->len);
    case OpKind::Binary:  // args: (lhs, rhs, out)
        return cb_binary(d->raw, op.dtype, op.binop, arg[0]->ptr, arg[1]->ptr, arg[2]->ptr, arg[2]->len);
    case OpKind::Apply2:  // args: (lhs, rhs, out): a fused two-input expression
        return cb_apply2(d->raw, op.expr, arg[0]->ptr, arg[1]->ptr, arg[2]->ptr, arg[2]->len);
    case OpKind::Clear:  // args: (buf)
        return cb_clear(d->raw, op.dtype, arg[0]->ptr, arg[0]->len);
    default: return CB_OK;
    }
}

// AddOperation::add_op: Lazy records (lazy.rs:93-107), everything else runs now (base.rs:53-62)
int32_t add_op(cbm_device *d, Op op)
{
    if (d->recording()) {
        // "each parent (id) must be unique": a check of Lazy's convert_to_operation only (lazy_graph.rs:112-128);
        // Base::add_op just calls the operation (base.rs:53-62) and the kernels allow out == in — after
        // optimize_mem_graph on Graph<Cached<..>> the buffers of a trace DO share one address
        for (size_t i = 0; i < op.arg_ids.size(); i++)
            for (size_t j = i + 1; j < op.arg_ids.size(); j++)
                if (op.arg_ids[i] == op.arg_ids[j])
                    return fail(CB_ERR_INVALID_ARG, "each parent (id) must be unique");
        d->ops.push_back(std::move(op));
        d->invalidate_replay();
        return CB_OK;
    }
    return call_op(d, op);
}

int32_t compile_one(cbm_device *d, int32_t dtype, int32_t kind, const cb_node *nodes, int32_t n, cb_expr **out)
{
    const cb_node *progs[1] = {nodes};
    const int32_t counts[1] = {n};
    return cb_expr_compile(d->raw, dtype, kind, progs, counts, 1, out);
}

// Retriever::retrieve through the module stack
int32_t retrieve(cbm_device *d, int32_t dtype, size_t len, const cbm_buf *parents, int32_t n_parents, cbm_buf *out)
{
    if (!valid_dtype(dtype)) return fail(CB_ERR_INVALID_ARG, "invalid dtype %d", dtype);
    if (len == 0) return fail(CB_ERR_ZERO_LENGTH, "retrieve: zero length buffer");
    std::vector<uint64_t> parent_ids;
    bool any_requires_grad = false;
    size_t global_len = 0, shard_begin = 0;
    for (int32_t i = 0; i < n_parents; i++) {
        GET_HANDLE(p, d, parents[i]);
        parent_ids.push_back(p->id);
        if (p->global_len && p->len == len) {  // an element-wise result of a slice is the same slice of the result
            global_len = p->global_len;
            shard_begin = p->shard_begin;
        }
        auto it = d->requires_grad.find(p->id);
        any_requires_grad = any_requires_grad || (it != d->requires_grad.end() && it->second);
    }

    Handle h;
    h.len = len;
    h.dtype = dtype;
    h.global_len = global_len;
    h.shard_begin = shard_begin;
    const uint64_t used_cursor = d->cursor;
    if (d->has(CBM_LAZY)) {
        // Lazy::retrieve: no allocation, the id is the cursor (lazy.rs:325-387)
        h.id = d->cursor;
        h.lazy = true;
        d->alloc_later.push_back(Deferred{h.id, len, dtype});
        d->cursor++;
    } else if (d->has(CBM_CACHED)) {
        // CachedModule::retrieve_entry (cached.rs:173-229)
        auto it = d->cache.find(d->cursor);
        if (it == d->cache.end()) {
            uint64_t p = 0;
            CB_TRY(alloc_zeroed(d, dtype, len, &p));
            it = d->cache.emplace(d->cursor, CacheSlot{p, len, dtype}).first;
        } else if (it->second.len * dtype_size(it->second.dtype) < len * dtype_size(dtype)) {
            return fail(CB_ERR_SHAPE, "cached buffer at cursor %llu is smaller than the requested one",
                        (unsigned long long)d->cursor);
        }
        h.ptr = it->second.ptr;
        h.id = h.ptr;
        d->cursor++;
        d->buffers[h.id] = Entry{h.ptr, len, dtype};
    } else {
        // Base::retrieve = device.alloc (base.rs:118-129); zeroed like the CPU device
        CB_TRY(alloc_zeroed(d, dtype, len, &h.ptr));
        h.id = h.ptr;
        h.owned = true;
        d->buffers[h.id] = Entry{h.ptr, len, dtype};
    }

    if (d->has(CBM_GRAPH) && !d->contains_ids.count(used_cursor)) {
        // Graph::retrieve_inner (graph.rs:159-195)
        d->contains_ids.insert(used_cursor);
        std::vector<size_t> deps;
        for (uint64_t pid : parent_ids) {
            auto it = d->buf_id_to_idx.find(pid);
            if (it == d->buf_id_to_idx.end())
                return fail(CB_ERR_GRAPH_OPTIMIZATION, "parent buffer id %llu is unknown to the graph", (unsigned long long)pid);
            deps.push_back(it->second);
        }
        const size_t idx = d->graph.size();
        d->buf_id_to_idx[h.id] = idx;
        d->idx_to_buf_id[idx] = h.id;
        d->idx_to_cursor[idx] = used_cursor;
        d->graph.add_node(len, std::move(deps));
    }
    if (d->has(CBM_AUTOGRAD)) d->requires_grad[h.id] = any_requires_grad;  // autograd.rs:122-128
    *out = d->new_handle(h);
    return CB_OK;
}

// Gradients::get_mut / get_ref: allocate (zeroed) on first use (gradients.rs:94-124, borrow_cache.rs:50-78)
int32_t grad_entry(cbm_device *d, uint64_t id, size_t len, int32_t dtype, Entry **out)
{
    auto it = d->grads.find(id);
    if (it == d->grads.end()) {
        uint64_t p = 0;
        CB_TRY(alloc_zeroed(d, dtype, len, &p));
        it = d->grads.emplace(id, Entry{p, len, dtype}).first;
    }
    *out = &it->second;
    return CB_OK;
}

int32_t run_ops(cbm_device *d)
{
    for (const Op &op : d->ops) CB_TRY(call_op(d, op));
    return CB_OK;
}

void free_replay(cbm_device *d)
{
    if (d->replay) cb_graph_destroy(d->replay);
    d->replay = nullptr;
    d->replay_valid = false;
}


// Recorded operations that READ buffer `id` (every argument an op does not write).  Ops added through
// cbm_binary_into / cbm_add_unary_grad / cbm_clear_op never went through retrieve(), so the graph has no edge for them:
// the fusing and aliasing passes count them here before they stop writing, or overwrite, a buffer.
size_t readers_of(const cbm_device *d, uint64_t id)
{
    size_t n = 0;
    for (const Op &o : d->ops) {
        if (o.kind == OpKind::NoOp) continue;
        for (size_t a = 0; a < o.arg_ids.size(); a++) {
            if (o.arg_ids[a] != id) continue;
            bool writes = false;
            switch (o.kind) {
            case OpKind::Apply: writes = a == 0; break;                 // (out, in)
            case OpKind::UnaryGrad: writes = false; break;              // lhs_grad is read-modify-write: counts as a reader
            case OpKind::Binary: case OpKind::Apply2: writes = a == 2; break;  // (lhs, rhs, out)
            case OpKind::Clear: writes = true; break;
            default: break;
            }
            if (!writes) n++;
        }
    }
    return n;
}

// What a pass that destroys the intermediates x_1 .. x_{K-1} of a unary chain ids = [x_0, x_1, .., x_K] (unary
// fusing: never written; memory-graph aliasing: overwritten) has to know about the Autograd tape.
enum class TapePlan { NoTape, Fused, Blocked };

// NoTape : no grad function touches an intermediate — nothing to do.
// Fused  : the chain's K grad functions (buf = x_{k-1}, out = x_k) sit on the tape next to each other and nothing else
//          touches an intermediate: they are replaced by ONE chain-grad entry (buf = x_0, out = x_K) that recomputes
//          the intermediates from x_0 (CB_KERNEL_CHAIN_GRAD).
// Blocked: the tape needs an intermediate in a way one kernel cannot replace — the caller must leave the chain alone.
int32_t tape_plan_for_chain(cbm_device *d, const std::vector<uint64_t> &ids, TapePlan *plan)
{
    *plan = TapePlan::NoTape;
    if (!d->has(CBM_AUTOGRAD) || d->tape.empty() || ids.size() < 3) return CB_OK;
    const size_t K = ids.size() - 1;
    auto is_intermediate = [&](uint64_t id) { return std::find(ids.begin() + 1, ids.end() - 1, id) != ids.end() - 1; };
    std::vector<size_t> touching;
    for (size_t p = 0; p < d->tape.size(); p++)
        if (is_intermediate(d->tape[p].buf_id) || is_intermediate(d->tape[p].out_id)) touching.push_back(p);
    if (touching.empty()) return CB_OK;
    *plan = TapePlan::Blocked;
    // the K entries must be tape[p0 .. p0+K), in chain order; the first of them is (x_0 -> x_1), which touches x_1
    const size_t p0 = touching.front();
    if (p0 + K > d->tape.size() || touching.size() != K || touching.back() != p0 + K - 1) return CB_OK;
    const int32_t dtype = d->tape[p0].dtype;
    for (size_t k = 0; k < K; k++) {
        const GradOp &g = d->tape[p0 + k];
        if (g.chain || g.buf_id != ids[k] || g.out_id != ids[k + 1] || g.dtype != dtype || g.fwd_ir.empty() || g.grad_ir.empty())
            return CB_OK;
    }
    // one flag for the whole chain: backward() tests requires_grad of x_0 only
    auto flag = [&](uint64_t id) {
        auto it = d->requires_grad.find(id);
        return it != d->requires_grad.end() && it->second;
    };
    for (size_t k = 1; k < K; k++)
        if (flag(ids[k]) != flag(ids[0])) return CB_OK;
    std::vector<const cb_node *> progs;
    std::vector<int32_t> counts;
    for (size_t k = 0; k < K; k++) {
        progs.push_back(d->tape[p0 + k].fwd_ir.data());
        counts.push_back((int32_t)d->tape[p0 + k].fwd_ir.size());
    }
    for (size_t k = 0; k < K; k++) {
        progs.push_back(d->tape[p0 + k].grad_ir.data());
        counts.push_back((int32_t)d->tape[p0 + k].grad_ir.size());
    }
    GradOp fused;
    fused.buf_id = ids[0];
    fused.out_id = ids[K];
    fused.dtype = dtype;
    fused.chain = true;
    fused.chain_len = K;
    CB_TRY(cb_expr_compile(d->raw, dtype, CB_KERNEL_CHAIN_GRAD, progs.data(), counts.data(), (int32_t)progs.size(), &fused.grad_expr));
    d->tape.erase(d->tape.begin() + (ptrdiff_t)p0 + 1, d->tape.begin() + (ptrdiff_t)(p0 + K));
    d->tape[p0] = std::move(fused);
    *plan = TapePlan::Fused;
    return CB_OK;
}

}  // namespace

// ===================================================================== device
extern "C" int32_t cbm_device_create(int32_t ordinal, uint32_t modules, int32_t dtype, cbm_device **out)
{
    CB_CHECK_ARG(out, "out is null");
    *out = nullptr;
    if (!valid_dtype(dtype)) return fail(CB_ERR_INVALID_ARG, "invalid dtype %d", dtype);
    if (modules & ~(CBM_CACHED | CBM_LAZY | CBM_GRAPH | CBM_AUTOGRAD)) return fail(CB_ERR_INVALID_ARG, "unknown module bits");
    if ((modules & CBM_GRAPH) && !(modules & (CBM_LAZY | CBM_CACHED)))
        return fail(CB_ERR_INVALID_ARG, "Graph needs a Cursor: stack it on Lazy or Cached (graph.rs:159-176)");
    std::unique_ptr<cbm_device> d(new cbm_device());
    d->mods = modules;
    d->dtype = dtype;
    CB_TRY(cb_device_create(ordinal, &d->raw));
    *out = d.release();
    return CB_OK;
}

extern "C" int32_t cbm_device_destroy(cbm_device *d)
{
    if (!d) return CB_OK;
    free_replay(d);
    for (auto &kv : d->handles)
        if (kv.second.owned && kv.second.ptr) cb_free(d->raw, kv.second.ptr);
    for (uint64_t p : d->lazy_allocs) cb_free(d->raw, p);
    {
        // after optimize_mem_graph several cursor positions share one allocation: free each once
        std::unordered_set<uint64_t> freed;
        for (auto &kv : d->cache)
            if (freed.insert(kv.second.ptr).second) cb_free(d->raw, kv.second.ptr);
        for (uint64_t p : d->cache_orphans)
            if (freed.insert(p).second) cb_free(d->raw, p);
    }
    for (auto &kv : d->grads) cb_free(d->raw, kv.second.ptr);
    cb_device_destroy(d->raw);
    delete d;
    return CB_OK;
}

extern "C" int32_t cbm_device_raw(cbm_device *d, cb_device **raw)
{
    CB_CHECK_ARG(d && raw, "null argument");
    *raw = d->raw;
    return CB_OK;
}

// ---- sharded device: one process per GPU, every buffer a contiguous slice ------------------------------------------
// The reference is one device per `CUDA::new` (src/devices/cuda/cuda.rs:53-67; its cross-device test is ignored,
// :187-202), so this is new.  Element-wise ops, fused chains and gradients are independent per element: each rank
// runs the unchanged operator API on its slice and never talks to a peer.  Only sum / mean combine — through the
// communicator's fused reduce + exchange kernel, with identical bits on every rank.
extern "C" int32_t cbm_device_set_comm(cbm_device *d, cb_comm *comm)
{
    CB_CHECK_ARG(d, "null device");
    if (comm) {
        cb_device *cd = nullptr;
        CB_TRY(cb_comm_device(comm, &cd));
        if (cd != d->raw) return fail(CB_ERR_INVALID_ARG, "the communicator belongs to another device (create it on cbm_device_raw)");
    }
    d->comm = comm;
    return CB_OK;
}

static int32_t shard_of(cbm_device *d, int32_t dtype, size_t global_len, size_t *begin, size_t *end)
{
    if (!d->comm) return fail(CB_ERR_STATE, "not a sharded device: call cbm_device_set_comm first");
    if (!valid_dtype(dtype)) return fail(CB_ERR_INVALID_ARG, "invalid dtype %d", dtype);
    int32_t rank = 0, n_ranks = 1;
    CB_TRY(cb_comm_rank(d->comm, &rank, &n_ranks));
    CB_TRY(cb_shard_range(global_len, (int32_t)dtype_size(dtype), n_ranks, rank, begin, end));
    if (*end <= *begin)
        return fail(CB_ERR_ZERO_LENGTH, "a buffer of %zu elements leaves rank %d of %d without a slice", global_len, rank, n_ranks);
    return CB_OK;
}

extern "C" int32_t cbm_buffer_new_sharded(cbm_device *d, int32_t dtype, size_t global_len, cbm_buf *out)
{
    CB_CHECK_ARG(d && out, "null argument");
    size_t b = 0, e = 0;
    CB_TRY(shard_of(d, dtype, global_len, &b, &e));
    CB_TRY(cbm_buffer_new(d, dtype, e - b, out));
    Handle *h = d->handle(*out);
    h->global_len = global_len;
    h->shard_begin = b;
    return CB_OK;
}

extern "C" int32_t cbm_buffer_from_host_sharded(cbm_device *d, int32_t dtype, const void *global_data, size_t global_len,
                                                cbm_buf *out)
{
    CB_CHECK_ARG(d && out && global_data, "null argument");
    size_t b = 0, e = 0;
    CB_TRY(shard_of(d, dtype, global_len, &b, &e));
    CB_TRY(cbm_buffer_from_host(d, dtype, static_cast<const char *>(global_data) + b * dtype_size(dtype), e - b, out));
    Handle *h = d->handle(*out);
    h->global_len = global_len;
    h->shard_begin = b;
    return CB_OK;
}

extern "C" int32_t cbm_buffer_shard(cbm_device *d, cbm_buf b, size_t *begin, size_t *end, size_t *global_len)
{
    CB_CHECK_ARG(d, "null device");
    GET_HANDLE(h, d, b);
    if (begin) *begin = h->global_len ? h->shard_begin : 0;
    if (end) *end = (h->global_len ? h->shard_begin : 0) + h->len;
    if (global_len) *global_len = h->global_len ? h->global_len : h->len;
    return CB_OK;
}

// ===================================================================== buffers
extern "C" int32_t cbm_buffer_new(cbm_device *d, int32_t dtype, size_t len, cbm_buf *out)
{
    CB_CHECK_ARG(d && out, "null argument");
    if (!valid_dtype(dtype)) return fail(CB_ERR_INVALID_ARG, "invalid dtype %d", dtype);
    Handle h;
    h.len = len;
    h.dtype = dtype;
    h.owned = true;
    CB_TRY(alloc_zeroed(d, dtype, len, &h.ptr));  // Buffer::new allocates at once, also under Lazy
    h.id = h.ptr;
    on_new_buffer(d, h);
    *out = d->new_handle(h);
    return CB_OK;
}

extern "C" int32_t cbm_buffer_from_host(cbm_device *d, int32_t dtype, const void *data, size_t len, cbm_buf *out)
{
    CB_CHECK_ARG(d && out && data, "null argument");
    if (!valid_dtype(dtype)) return fail(CB_ERR_INVALID_ARG, "invalid dtype %d", dtype);
    if (len == 0) return fail(CB_ERR_ZERO_LENGTH, "zero length buffer");
    Handle h;
    h.len = len;
    h.dtype = dtype;
    h.owned = true;
    CB_TRY(cb_alloc(d->raw, len * dtype_size(dtype), 0, &h.ptr));  // alloc_from_slice (cuda.rs:124-137)
    int32_t rc = cb_h2d(d->raw, h.ptr, data, len * dtype_size(dtype));
    if (rc != CB_OK) {
        cb_free(d->raw, h.ptr);
        return rc;
    }
    h.id = h.ptr;
    on_new_buffer(d, h);
    *out = d->new_handle(h);
    return CB_OK;
}

extern "C" int32_t cbm_buffer_drop(cbm_device *d, cbm_buf b)
{
    CB_CHECK_ARG(d, "null device");
    GET_HANDLE(h, d, b);
    // the registered shallow copy goes away with the buffer: recorded ops that still name the
    // id fail with InvalidLazyBuf at run() (lazy.rs:624-640)
    bool is_grad_view = false;
    for (auto &kv : d->grad_handle) is_grad_view = is_grad_view || kv.second == b;
    const bool last_ref = --d->id_refs[h->id] <= 0;
    if (last_ref) d->id_refs.erase(h->id);
    if (!is_grad_view && last_ref) d->buffers.erase(h->id);
    if (h->owned && h->ptr && last_ref && d->has(CBM_AUTOGRAD)) {
        // The id of an eager buffer is its address, and the pool hands a freed address straight back: state keyed by
        // the id must not outlive the buffer, or the next buffer at that address inherits requires_grad and the old
        // accumulated gradient.  (The reference never removes it — ids there are addresses too, autograd.rs:86-91.)
        const uint64_t id = h->id;
        d->requires_grad.erase(id);
        auto gh = d->grad_handle.find(id);
        if (gh != d->grad_handle.end()) {
            Handle *view = d->handle(gh->second);
            if (view) {
                if (--d->id_refs[view->id] <= 0) d->id_refs.erase(view->id);
                d->buffers.erase(view->id);
                d->handles.erase(gh->second);  // the lent view dies with the gradient it points to
            }
            d->grad_handle.erase(gh);
        }
        auto g = d->grads.find(id);
        if (g != d->grads.end()) {
            d->buffers.erase(g->second.ptr);
            cb_free(d->raw, g->second.ptr);
            d->grads.erase(g);
        }
    }
    if (h->owned && h->ptr) CB_TRY(cb_free(d->raw, h->ptr));
    d->handles.erase(b);
    d->invalidate_replay();
    return CB_OK;
}

extern "C" int32_t cbm_buffer_len(cbm_device *d, cbm_buf b, size_t *len)
{
    CB_CHECK_ARG(d && len, "null argument");
    GET_HANDLE(h, d, b);
    *len = h->len;
    return CB_OK;
}

extern "C" int32_t cbm_buffer_dtype(cbm_device *d, cbm_buf b, int32_t *dtype)
{
    CB_CHECK_ARG(d && dtype, "null argument");
    GET_HANDLE(h, d, b);
    *dtype = h->dtype;
    return CB_OK;
}

static int32_t storage_of(cbm_device *d, const Handle *h, Entry *out)
{
    if (!h->lazy && h->ptr && !d->resolve(h->id)) {  // gradient views and the like
        *out = Entry{h->ptr, h->len, h->dtype};
        return CB_OK;
    }
    const Entry *e = d->resolve(h->id);
    if (!e) return fail(CB_ERR_INVALID_LAZY_BUF, "buffer id %llu has no storage yet (lazy buffer before run()/alloc_later())",
                        (unsigned long long)h->id);
    *out = *e;
    return CB_OK;
}

extern "C" int32_t cbm_buffer_read(cbm_device *d, cbm_buf b, void *host_out, size_t len)
{
    CB_CHECK_ARG(d && host_out, "null argument");
    GET_HANDLE(h, d, b);
    Entry e;
    CB_TRY(storage_of(d, h, &e));
    if (len > e.len) return fail(CB_ERR_SHAPE, "read of %zu elements from a buffer of %zu", len, e.len);
    return cb_d2h(d->raw, host_out, e.ptr, len * dtype_size(h->dtype));
}

extern "C" int32_t cbm_buffer_write(cbm_device *d, cbm_buf b, const void *host_in, size_t len)
{
    CB_CHECK_ARG(d && host_in, "null argument");
    GET_HANDLE(h, d, b);
    Entry e;
    CB_TRY(storage_of(d, h, &e));
    if (len > e.len) return fail(CB_ERR_SHAPE, "write of %zu elements into a buffer of %zu", len, e.len);
    return cb_h2d(d->raw, e.ptr, host_in, len * dtype_size(h->dtype));
}

extern "C" int32_t cbm_buffer_ptr(cbm_device *d, cbm_buf b, uint64_t *dptr)
{
    CB_CHECK_ARG(d && dptr, "null argument");
    GET_HANDLE(h, d, b);
    const Entry *e = d->resolve(h->id);
    *dptr = e ? e->ptr : h->ptr;
    return CB_OK;
}

extern "C" int32_t cbm_buffer_id(cbm_device *d, cbm_buf b, uint64_t *id)
{
    CB_CHECK_ARG(d && id, "null argument");
    GET_HANDLE(h, d, b);
    *id = h->id;
    return CB_OK;
}

extern "C" int32_t cbm_buffer_require_grad(cbm_device *d, cbm_buf b)
{
    CB_CHECK_ARG(d, "null device");
    GET_HANDLE(h, d, b);
    d->requires_grad[h->id] = true;
    return CB_OK;
}

extern "C" int32_t cbm_buffer_requires_grad(cbm_device *d, cbm_buf b, int32_t *flag)
{
    CB_CHECK_ARG(d && flag, "null argument");
    GET_HANDLE(h, d, b);
    auto it = d->requires_grad.find(h->id);
    *flag = (it != d->requires_grad.end() && it->second) ? 1 : 0;
    return CB_OK;
}

extern "C" int32_t cbm_buffer_checkpoint(cbm_device *d, cbm_buf b)
{
    CB_CHECK_ARG(d, "null device");
    GET_HANDLE(h, d, b);
    if (!d->has(CBM_GRAPH)) return CB_OK;  // set_checkpoint_buffer passes down to nothing
    auto it = d->buf_id_to_idx.find(h->id);
    if (it == d->buf_id_to_idx.end()) return fail(CB_ERR_GRAPH_OPTIMIZATION, "buffer is unknown to the graph");
    d->graph.node(it->second).skip = true;  // graph.rs:255-259
    return CB_OK;
}

// ===================================================================== operations
extern "C" int32_t cbm_retrieve(cbm_device *d, int32_t dtype, size_t len, const cbm_buf *parents, int32_t n_parents,
                                cbm_buf *out)
{
    CB_CHECK_ARG(d && out && (n_parents == 0 || parents) && n_parents >= 0, "bad argument");
    return retrieve(d, dtype, len, parents, n_parents, out);
}

extern "C" int32_t cbm_apply_fn(cbm_device *d, cbm_buf in, const cb_node *nodes, int32_t n_nodes, cbm_buf *out)
{
    CB_CHECK_ARG(d && out, "null argument");
    GET_HANDLE(hin, d, in);
    const int32_t dtype = hin->dtype;
    const uint64_t in_id = hin->id;
    const size_t len = hin->len;
    cb_expr *e = nullptr;
    CB_TRY(compile_one(d, dtype, CB_KERNEL_APPLY, nodes, n_nodes, &e));
    // let mut out = self.retrieve(buf.len(), buf); self.add_op((&mut out, buf), ..); self.set_op_hint(unary(f))
    // (src/devices/cuda/ops.rs:134-140)
    cbm_buf ob = 0;
    CB_TRY(retrieve(d, dtype, len, &in, 1, &ob));
    const uint64_t out_id = d->handle(ob)->id;
    Op op;
    op.kind = OpKind::Apply;
    op.arg_ids = {out_id, in_id};
    op.dtype = dtype;
    op.expr = e;
    if (d->recording()) {  // Lazy::set_op_hint (lazy.rs:136-143); Base ignores hints (base.rs:86)
        op.unary_hint = true;
        op.hint.assign(nodes, nodes + n_nodes);
        op.tree = op.hint;
        op.in = {in_id};
        op.out = out_id;
        d->op_of_id[out_id] = d->ops.size();
    }
    int32_t rc = add_op(d, std::move(op));
    if (rc != CB_OK) return rc;
    *out = ob;
    return CB_OK;
}

extern "C" int32_t cbm_add_unary_grad(cbm_device *d, cbm_buf lhs, cbm_buf lhs_grad, cbm_buf out_grad, const cb_node *nodes,
                                      int32_t n_nodes)
{
    CB_CHECK_ARG(d, "null device");
    GET_HANDLE(hl, d, lhs);
    GET_HANDLE(hg, d, lhs_grad);
    GET_HANDLE(ho, d, out_grad);
    if (hl->dtype != hg->dtype || hl->dtype != ho->dtype) return fail(CB_ERR_INVALID_ARG, "dtype mismatch");
    if (hg->len < hl->len || ho->len < hl->len) return fail(CB_ERR_SHAPE, "gradient buffers are shorter than lhs");
    cb_expr *e = nullptr;
    CB_TRY(compile_one(d, hl->dtype, CB_KERNEL_UNARY_GRAD, nodes, n_nodes, &e));
    Op op;
    op.kind = OpKind::UnaryGrad;
    op.arg_ids = {hl->id, hg->id, ho->id};  // (lhs, lhs_grad, out): src/devices/cuda/ops.rs:198
    op.dtype = hl->dtype;
    op.expr = e;
    return add_op(d, std::move(op));
}

extern "C" int32_t cbm_unary_ew(cbm_device *d, cbm_buf in, const cb_node *fwd, int32_t n_fwd, const cb_node *grad,
                                int32_t n_grad, cbm_buf *out)
{
    CB_CHECK_ARG(d && out, "null argument");
    GET_HANDLE(hin, d, in);
    const int32_t dtype = hin->dtype;
    const uint64_t in_id = hin->id;
    // the grad closure is validated up front even when it is never recorded
    CB_TRY(expr_validate(dtype, CB_KERNEL_UNARY_GRAD, grad, n_grad));
    CB_TRY(cbm_apply_fn(d, in, fwd, n_fwd, out));
    // add_grad_fn: a no-op without Autograd or with gradients disabled (autograd.rs:248-258)
    if (!d->has(CBM_AUTOGRAD) || !d->grad_enabled) return CB_OK;
    GradOp g;
    g.buf_id = in_id;
    g.out_id = d->handle(*out)->id;
    g.dtype = dtype;
    CB_TRY(compile_one(d, dtype, CB_KERNEL_UNARY_GRAD, grad, n_grad, &g.grad_expr));
    g.fwd_ir.assign(fwd, fwd + n_fwd);
    g.grad_ir.assign(grad, grad + n_grad);
    d->tape.push_back(std::move(g));
    return CB_OK;
}

extern "C" int32_t cbm_binary(cbm_device *d, int32_t op, cbm_buf lhs, cbm_buf rhs, cbm_buf *out)
{
    CB_CHECK_ARG(d && out, "null argument");
    CB_CHECK_ARG(op >= CB_BIN_ADD && op <= CB_BIN_DIV, "bad binary op");
    GET_HANDLE(hl, d, lhs);
    GET_HANDLE(hr, d, rhs);
    if (hl->dtype != hr->dtype) return fail(CB_ERR_INVALID_ARG, "dtype mismatch");
    if (hl->len != hr->len) return fail(CB_ERR_SHAPE, "length mismatch: %zu vs %zu", hl->len, hr->len);
    const uint64_t lid = hl->id, rid = hr->id;
    const int32_t dtype = hl->dtype;
    const size_t len = hl->len;
    if (lid == rid && d->recording()) return fail(CB_ERR_INVALID_ARG, "each parent (id) must be unique");  // lazy_graph.rs:112-128
    // README.md:96-122: retrieve(len, (lhs, rhs)) then add_op((lhs, rhs, &mut out), ..)
    const cbm_buf parents[2] = {lhs, rhs};
    cbm_buf ob = 0;
    CB_TRY(retrieve(d, dtype, len, parents, 2, &ob));
    const uint64_t out_id = d->handle(ob)->id;
    Op o;
    o.kind = OpKind::Binary;
    o.dtype = dtype;
    o.binop = op;
    o.arg_ids = {lid, rid, out_id};
    if (d->recording()) {
        d->op_of_id[out_id] = d->ops.size();
        o.in = {lid, rid};
        o.out = out_id;
        o.tree = {mk_node(CB_OP_X), mk_node(CB_OP_Y), mk_node(CB_OP_ADD + op, 0, 1)};
    }
    int32_t rc = add_op(d, std::move(o));
    if (rc != CB_OK) return rc;
    *out = ob;
    return CB_OK;
}

// A binary op whose output buffer is chosen by the caller — what the reference's tests do with their own
// kernels: `device.launch_kernel1d(len, &add_src, "add", &[&lhs, &rhs, &mut out, &len])`
// (src/devices/cuda/lazy.rs:96-141) or `add_op((&a, &b, &mut out), ..)` (src/modules/lazy.rs:741-751).
// No retrieve, hence no graph node and no cache trace: the fusing passes treat it as an opaque reader /
// writer of its three buffers.
extern "C" int32_t cbm_binary_into(cbm_device *d, int32_t op, cbm_buf lhs, cbm_buf rhs, cbm_buf out)
{
    CB_CHECK_ARG(d, "null device");
    CB_CHECK_ARG(op >= CB_BIN_ADD && op <= CB_BIN_DIV, "bad binary op");
    GET_HANDLE(hl, d, lhs);
    GET_HANDLE(hr, d, rhs);
    GET_HANDLE(ho, d, out);
    if (hl->dtype != hr->dtype || hl->dtype != ho->dtype) return fail(CB_ERR_INVALID_ARG, "dtype mismatch");
    if (hl->len != hr->len || hl->len != ho->len)
        return fail(CB_ERR_SHAPE, "length mismatch: %zu, %zu -> %zu", hl->len, hr->len, ho->len);
    Op o;
    o.kind = OpKind::Binary;
    o.dtype = hl->dtype;
    o.binop = op;
    o.arg_ids = {hl->id, hr->id, ho->id};
    o.in = {hl->id, hr->id};
    o.out = ho->id;
    return add_op(d, std::move(o));
}

// `add_op(&mut out, |out, _| { out.clear(); Ok(()) })` (src/modules/lazy.rs:733-738): a clear that is
// RECORDED under Lazy, unlike ClearBuf::clear below, which the reference runs eagerly on every stack.
extern "C" int32_t cbm_clear_op(cbm_device *d, cbm_buf b)
{
    CB_CHECK_ARG(d, "null device");
    GET_HANDLE(h, d, b);
    Op o;
    o.kind = OpKind::Clear;
    o.dtype = h->dtype;
    o.arg_ids = {h->id};
    o.out = h->id;
    return add_op(d, std::move(o));
}

extern "C" int32_t cbm_clear(cbm_device *d, cbm_buf b)
{
    CB_CHECK_ARG(d, "null device");
    GET_HANDLE(h, d, b);
    Entry e;
    CB_TRY(storage_of(d, h, &e));
    return cb_clear(d->raw, h->dtype, e.ptr, e.len);  // ClearBuf::clear is eager (src/devices/cuda/ops.rs:51-56)
}

extern "C" int32_t cbm_copy_slice(cbm_device *d, cbm_buf src, size_t src_off, cbm_buf dst, size_t dst_off, size_t n)
{
    CB_CHECK_ARG(d, "null device");
    GET_HANDLE(hs, d, src);
    GET_HANDLE(hd, d, dst);
    if (hs->dtype != hd->dtype) return fail(CB_ERR_INVALID_ARG, "dtype mismatch");
    Entry es, ed;
    CB_TRY(storage_of(d, hs, &es));
    CB_TRY(storage_of(d, hd, &ed));
    if (src_off + n > es.len || dst_off + n > ed.len) return fail(CB_ERR_SHAPE, "slice out of range");
    return cb_copy(d->raw, hs->dtype, ed.ptr, dst_off, es.ptr, src_off, n);
}

extern "C" int32_t cbm_clone_buf(cbm_device *d, cbm_buf src, cbm_buf *out)
{
    CB_CHECK_ARG(d && out, "null argument");
    GET_HANDLE(hs, d, src);
    Entry es;
    CB_TRY(storage_of(d, hs, &es));
    const int32_t dtype = hs->dtype;
    cbm_buf nb = 0;
    const size_t global_len = hs->global_len, shard_begin = hs->shard_begin;
    CB_TRY(cbm_buffer_new(d, dtype, es.len, &nb));  // CloneBuf (src/devices/cuda/cuda.rs:152-164)
    d->handle(nb)->global_len = global_len;
    d->handle(nb)->shard_begin = shard_begin;
    CB_TRY(cb_copy(d->raw, dtype, d->handle(nb)->ptr, 0, es.ptr, 0, es.len));
    *out = nb;
    return CB_OK;
}

extern "C" int32_t cbm_sum(cbm_device *d, cbm_buf b, void *host_out)
{
    CB_CHECK_ARG(d && host_out, "null argument");
    GET_HANDLE(h, d, b);
    Entry e;
    CB_TRY(storage_of(d, h, &e));
    if (d->comm && h->global_len)  // the rank's partial, then ONE exchange of a scalar per rank (rank-ordered fold)
        return cb_comm_sum_host(d->comm, h->dtype, e.ptr, e.len, host_out);
    return cb_sum_host(d->raw, h->dtype, e.ptr, e.len, host_out);
}

extern "C" int32_t cbm_mean(cbm_device *d, cbm_buf b, void *host_out)
{
    CB_CHECK_ARG(d && host_out, "null argument");
    GET_HANDLE(h, d, b);
    Entry e;
    CB_TRY(storage_of(d, h, &e));
    if (d->comm && h->global_len) return cb_comm_mean_host(d->comm, h->dtype, e.ptr, e.len, h->global_len, host_out);
    return cb_mean_host(d->raw, h->dtype, e.ptr, e.len, host_out);
}

// ===================================================================== Lazy
extern "C" int32_t cbm_alloc_later(cbm_device *d)
{
    CB_CHECK_ARG(d, "null device");
    return run_alloc_later(d);
}

extern "C" int32_t cbm_run(cbm_device *d)
{
    CB_CHECK_ARG(d, "null device");
    if (!d->has(CBM_LAZY)) return CB_OK;  // Base: RunModule is a no-op
    CB_TRY(run_alloc_later(d));           // lazy.rs:190-198: alloc, call every op, device.run()
    if (!d->replay_enabled) return run_ops(d);
    if (!d->replay_valid) {
        // capture once, after every buffer exists; afterwards run() is a single cudaGraphLaunch
        free_replay(d);
        for (const Op &op : d->ops)  // fail before the capture starts if a buffer went away
            for (uint64_t id : op.arg_ids)
                if (op.kind != OpKind::NoOp && !d->resolve(id))
                    return fail(CB_ERR_INVALID_LAZY_BUF, "InvalidLazyBuf: buffer id %llu is not alive", (unsigned long long)id);
        CB_TRY(cb_graph_begin(d->raw));
        int32_t rc = run_ops(d);
        cb_graph *g = nullptr;
        int32_t rc2 = cb_graph_end(d->raw, &g);
        if (rc != CB_OK) {
            if (g) cb_graph_destroy(g);
            return rc;
        }
        CB_TRY(rc2);
        d->replay = g;
        d->replay_valid = true;
    }
    return cb_graph_launch(d->raw, d->replay);
}

extern "C" int32_t cbm_exec_now(cbm_device *d, size_t begin, size_t end)
{
    CB_CHECK_ARG(d, "null device");
    if (!d->has(CBM_LAZY)) return CB_OK;
    CB_TRY(run_alloc_later(d));
    if (end > d->ops.size()) end = d->ops.size();
    if (begin > end) return fail(CB_ERR_INVALID_ARG, "exec_now: begin %zu > end %zu", begin, end);
    // call_range drains the operations it executes (lazy_graph.rs:92-104)
    std::vector<Op> taken(std::make_move_iterator(d->ops.begin() + (ptrdiff_t)begin),
                          std::make_move_iterator(d->ops.begin() + (ptrdiff_t)end));
    d->ops.erase(d->ops.begin() + (ptrdiff_t)begin, d->ops.begin() + (ptrdiff_t)end);
    d->op_of_id.clear();  // positions moved; fusing needs a freshly recorded sequence
    d->invalidate_replay();
    for (const Op &op : taken) CB_TRY(call_op(d, op));
    return CB_OK;
}

extern "C" int32_t cbm_exec_last_n(cbm_device *d, size_t n)
{
    CB_CHECK_ARG(d, "null device");
    const size_t total = d->ops.size();
    return cbm_exec_now(d, n > total ? 0 : total - n, total);  // features.rs:505-518
}

extern "C" int32_t cbm_ops_count(cbm_device *d, size_t *n)
{
    CB_CHECK_ARG(d && n, "null argument");
    *n = d->ops.size();
    return CB_OK;
}

extern "C" int32_t cbm_set_lazy_enabled(cbm_device *d, int32_t enabled)
{
    CB_CHECK_ARG(d, "null device");
    d->lazy_enabled = enabled != 0;
    return CB_OK;
}

extern "C" int32_t cbm_op_hint_src(cbm_device *d, size_t i, char *out, size_t cap)
{
    CB_CHECK_ARG(d && out && cap > 0, "bad argument");
    if (i >= d->ops.size()) return fail(CB_ERR_INVALID_ARG, "op index %zu out of range (%zu ops)", i, d->ops.size());
    const Op &op = d->ops[i];
    std::string s;
    if (op.unary_hint) s = expr_to_cl_source(op.dtype, op.hint.data(), (int32_t)op.hint.size(), "x", "y");
    else if (op.fused && op.kind == OpKind::Apply2) s = "Fused: " + expr_to_cl_source(op.dtype, op.tree.data(), (int32_t)op.tree.size(), "x", "y");
    else if (op.fused) s = "UnaryFused";
    if (s.size() + 1 > cap) return fail(CB_ERR_INVALID_ARG, "output buffer too small");
    std::memcpy(out, s.c_str(), s.size() + 1);
    return CB_OK;
}

// the compiled kernel behind recorded op i (after the fusing passes): lets a caller hand the SAME fused chain to
// cb_apply_host for host-resident operands
extern "C" int32_t cbm_op_expr(cbm_device *d, size_t i, cb_expr **out)
{
    CB_CHECK_ARG(d && out, "bad argument");
    *out = nullptr;
    if (i >= d->ops.size()) return fail(CB_ERR_INVALID_ARG, "op index %zu out of range (%zu ops)", i, d->ops.size());
    *out = d->ops[i].expr;  // null for no-ops and the AOT kernels (binary, clear)
    return CB_OK;
}

extern "C" int32_t cbm_set_graph_replay(cbm_device *d, int32_t enabled)
{
    CB_CHECK_ARG(d, "null device");
    d->replay_enabled = enabled != 0;
    if (!d->replay_enabled) free_replay(d);
    return CB_OK;
}

extern "C" int32_t cbm_replay_kernel_nodes(cbm_device *d, size_t *n)
{
    CB_CHECK_ARG(d && n, "null argument");
    *n = 0;
    if (d->replay && d->replay_valid) return cb_graph_node_count(d->replay, n);
    return CB_OK;
}

// ===================================================================== Graph
extern "C" int32_t cbm_cache_traces(cbm_device *d, int64_t *out, size_t cap, size_t *written)
{
    CB_CHECK_ARG(d, "null device");
    if (!d->has(CBM_GRAPH)) return fail(CB_ERR_MISSING_CACHE_TRACES, "MissingCacheTraces: no Graph module");
    return cb_flatten_traces(d->graph.cache_traces(), out, cap, written);
}

extern "C" int32_t cbm_optimize_mem_graph(cbm_device *d)
{
    CB_CHECK_ARG(d, "null device");
    if (!d->has(CBM_GRAPH)) return fail(CB_ERR_MISSING_CACHE_TRACES, "MissingCacheTraces: no Graph module");
    const std::vector<CacheTrace> traces = d->graph.cache_traces();
    if (d->has(CBM_LAZY)) {
        // Lazy::alloc_later_optimized (lazy/optimization.rs:4-44): allocate the head of each trace
        // and register the same storage under every id of the trace
        std::unordered_set<uint64_t> aliased;
        std::vector<Deferred> pending;
        pending.swap(d->alloc_later);
        std::unordered_map<uint64_t, Deferred> by_id;
        for (const Deferred &a : pending) by_id.emplace(a.id, a);
        for (const CacheTrace &t : traces) {
            auto head = d->idx_to_buf_id.find(t.cache_idx);
            if (head == d->idx_to_buf_id.end()) return fail(CB_ERR_GRAPH_OPTIMIZATION, "GraphOptimization: trace head has no buffer");
            auto def = by_id.find(head->second);
            if (def == by_id.end()) continue;  // allocated earlier: nothing to share any more
            const Deferred a = def->second;
            {
                // every member but the last is overwritten by its successor: it may have that one reader only
                bool extra_reader = readers_of(d, head->second) > 1;
                for (size_t m = 0; m + 1 < t.use_cache_idxs.size() && !extra_reader; m++) {
                    auto uid = d->idx_to_buf_id.find(t.use_cache_idxs[m]);
                    extra_reader = uid != d->idx_to_buf_id.end() && readers_of(d, uid->second) > 1;
                }
                if (extra_reader) continue;
            }
            if (d->has(CBM_AUTOGRAD) && !d->tape.empty()) {
                // sharing one allocation overwrites every buffer of the trace but the last; grad functions that read
                // them must become one recomputing chain-grad kernel first, else the trace keeps its own buffers
                std::vector<uint64_t> chain_ids;
                auto prod = d->op_of_id.find(head->second);
                if (prod != d->op_of_id.end() && prod->second < d->ops.size() && d->ops[prod->second].kind == OpKind::Apply &&
                    d->ops[prod->second].arg_ids.size() == 2)
                    chain_ids.push_back(d->ops[prod->second].arg_ids[1]);
                else
                    chain_ids.push_back(UINT64_MAX);  // no unary producer: an id no grad function can name
                chain_ids.push_back(head->second);
                for (size_t use : t.use_cache_idxs) {
                    auto uid = d->idx_to_buf_id.find(use);
                    if (uid != d->idx_to_buf_id.end()) chain_ids.push_back(uid->second);
                }
                // the head is overwritten too, so it counts as an intermediate: prepend the chain's real input
                TapePlan plan = TapePlan::NoTape;
                CB_TRY(tape_plan_for_chain(d, chain_ids, &plan));
                if (plan == TapePlan::Blocked) continue;
            }
            if (!d->allocated_ids.count(a.id)) {
                if (d->buffers.count(a.id)) return fail(CB_ERR_STATE, "IDs collided! Maybe pointing address already occupied this ID.");
                uint64_t p = 0;
                CB_TRY(alloc_zeroed(d, a.dtype, a.len, &p));
                d->lazy_allocs.push_back(p);
                d->allocated_ids.insert(a.id);
                d->buffers[a.id] = Entry{p, a.len, a.dtype};
            }
            aliased.insert(a.id);
            const Entry shared = d->buffers[a.id];
            for (size_t use : t.use_cache_idxs) {
                auto uid = d->idx_to_buf_id.find(use);
                if (uid == d->idx_to_buf_id.end()) return fail(CB_ERR_GRAPH_OPTIMIZATION, "GraphOptimization: trace node has no buffer");
                auto udef = by_id.find(uid->second);
                if (udef == by_id.end()) continue;
                if (udef->second.dtype != a.dtype || udef->second.len != a.len) continue;  // never alias across types
                d->buffers[uid->second] = shared;
                d->allocated_ids.insert(uid->second);
                aliased.insert(uid->second);
            }
        }
        for (const Deferred &a : pending)
            if (!aliased.count(a.id)) d->alloc_later.push_back(a);  // not on a trace: allocated by run()
        d->invalidate_replay();
        return CB_OK;
    }
    // Cached::optimize_mem_graph (cached.rs:414-448): later iterations of the loop hand out the
    // head's allocation for every cursor position of the trace
    for (const CacheTrace &t : traces) {
        if (d->has(CBM_AUTOGRAD)) {
            // eager stacks re-record the tape every iteration; a trace whose buffers carry gradients keeps its own
            // allocations (their values are what the next backward() reads)
            bool needs_values = false;
            std::vector<size_t> members{t.cache_idx};
            members.insert(members.end(), t.use_cache_idxs.begin(), t.use_cache_idxs.end());
            for (size_t m : members) {
                auto id = d->idx_to_buf_id.find(m);
                if (id == d->idx_to_buf_id.end()) continue;
                auto rg = d->requires_grad.find(id->second);
                needs_values = needs_values || (rg != d->requires_grad.end() && rg->second);
            }
            if (needs_values) continue;
        }
        auto hc = d->idx_to_cursor.find(t.cache_idx);
        if (hc == d->idx_to_cursor.end()) return fail(CB_ERR_GRAPH_OPTIMIZATION, "GraphOptimization: trace head has no cursor");
        auto head = d->cache.find(hc->second);
        if (head == d->cache.end()) return fail(CB_ERR_GRAPH_OPTIMIZATION, "GraphOptimization: trace head is not cached");
        for (size_t use : t.use_cache_idxs) {
            auto uc = d->idx_to_cursor.find(use);
            if (uc == d->idx_to_cursor.end()) continue;
            auto slot = d->cache.find(uc->second);
            if (slot == d->cache.end()) continue;
            if (slot->second.dtype != head->second.dtype || slot->second.len != head->second.len) continue;
            if (slot->second.ptr != head->second.ptr) d->cache_orphans.push_back(slot->second.ptr);  // still used by this iteration's handles
            slot->second = head->second;
        }
    }
    return CB_OK;
}

extern "C" int32_t cbm_unary_fusing(cbm_device *d)
{
    CB_CHECK_ARG(d, "null device");
    if (!d->has(CBM_GRAPH)) return fail(CB_ERR_MISSING_CACHE_TRACES, "MissingCacheTraces: no Graph module");
    if (!d->has(CBM_LAZY)) return fail(CB_ERR_UNSUPPORTED, "UnaryFusingUnsupported: only Lazy records operations");
    // Lazy::fuse_unary_ops (lazy/optimization.rs:46-95) + UnaryFusing::fuse_unary_ops (devices/fusing.rs:41-92)
    for (const CacheTrace &t : d->graph.cache_traces()) {
        std::vector<size_t> idxs{t.cache_idx};
        idxs.insert(idxs.end(), t.use_cache_idxs.begin(), t.use_cache_idxs.end());
        // node index -> id of the retrieved buffer -> the op that writes it
        std::vector<long> op_idx;
        for (size_t idx : idxs) {
            long oi = -1;
            auto id = d->idx_to_buf_id.find(idx);
            if (id != d->idx_to_buf_id.end()) {
                auto o = d->op_of_id.find(id->second);
                if (o != d->op_of_id.end() && o->second < d->ops.size()) oi = (long)o->second;
            }
            op_idx.push_back(oi);
        }
        size_t i = 0;
        auto on_tape = [&](const Op &o) {
            for (const GradOp &g : d->tape)
                if (!g.chain && g.buf_id == o.arg_ids[1] && g.out_id == o.arg_ids[0]) return true;
            return false;
        };
        while (i < op_idx.size()) {
            auto fusable = [&](size_t k, int32_t dtype, size_t prev_k) {
                if (op_idx[k] < 0) return false;
                const Op &o = d->ops[(size_t)op_idx[k]];
                if (!(o.kind == OpKind::Apply && o.unary_hint) || o.dtype != dtype) return false;
                if (k == prev_k) return true;
                // the op must consume what the previous op of the run produced, and be its ONLY reader: an op recorded
                // with a caller-owned output (cbm_binary_into, a recorded add_unary_grad) reads buffers the graph knows
                // nothing about, and fusing the producer away would feed it zeros
                const uint64_t mid = d->ops[(size_t)op_idx[prev_k]].arg_ids[0];
                if (o.arg_ids[1] != mid || readers_of(d, mid) != 1) return false;
                // a run is made of ops that all have a grad function on the tape (unary_ew) or that all have none
                // (apply_fn): a mixed chain is split where that changes, so each part can be fused — the recorded part
                // together with its grad functions — instead of the whole chain staying unfused
                return on_tape(o) == on_tape(d->ops[(size_t)op_idx[prev_k]]);
            };
            if (op_idx[i] < 0 || !(d->ops[(size_t)op_idx[i]].kind == OpKind::Apply && d->ops[(size_t)op_idx[i]].unary_hint)) {
                i++;
                continue;
            }
            const int32_t dtype = d->ops[(size_t)op_idx[i]].dtype;
            size_t j = i + 1;
            while (j < op_idx.size() && fusable(j, dtype, j - 1)) j++;
            TapePlan plan = TapePlan::NoTape;
            if (j - i >= 2) {
                // the grad functions of unary_ew read the buffers this run would stop writing: fuse them as well
                // (one recomputing chain-grad kernel), or leave the run alone when that is not possible
                std::vector<uint64_t> chain_ids{d->ops[(size_t)op_idx[i]].arg_ids[1]};
                for (size_t k = i; k < j; k++) chain_ids.push_back(d->ops[(size_t)op_idx[k]].arg_ids[0]);
                CB_TRY(tape_plan_for_chain(d, chain_ids, &plan));
            }
            if (j - i >= 2 && plan != TapePlan::Blocked) {
                Op &first = d->ops[(size_t)op_idx[i]];
                const Op &last = d->ops[(size_t)op_idx[j - 1]];
                // out = the last op's output, in = the first op's input; they must differ (fusing.rs:60-67)
                if (last.arg_ids[0] == first.arg_ids[1]) return fail(CB_ERR_STATE, "fused chain would read and write the same buffer id");
                std::vector<const cb_node *> progs;
                std::vector<int32_t> counts;
                for (size_t k = i; k < j; k++) {
                    const Op &o = d->ops[(size_t)op_idx[k]];
                    progs.push_back(o.hint.data());
                    counts.push_back((int32_t)o.hint.size());
                }
                cb_expr *fused = nullptr;
                CB_TRY(cb_expr_compile(d->raw, dtype, CB_KERNEL_APPLY, progs.data(), counts.data(), (int32_t)progs.size(), &fused));
                const uint64_t out_id = last.arg_ids[0], in_id = first.arg_ids[1];
                std::vector<cb_node> tree = d->ops[(size_t)op_idx[i]].tree;
                const int32_t xmap[2] = {CB_OP_X, CB_OP_X};
                for (size_t k = i + 1; k < j; k++) tree = substitute(d->ops[(size_t)op_idx[k]].tree, CB_OP_X, CB_OP_X, tree, xmap);
                for (size_t k = i + 1; k < j; k++) d->ops[(size_t)op_idx[k]] = Op();
                first.tree = std::move(tree);
                first.in = {in_id};
                first.out = out_id;
                first.expr = fused;
                first.arg_ids = {out_id, in_id};
                first.unary_hint = false;
                first.hint.clear();
                first.fused = true;  // OpHint::UnaryFused
            }
            i = j;
        }
    }
    d->invalidate_replay();
    return CB_OK;
}

// Element-wise fusing (not in the reference, which only fuses unary chains):
// a producer op whose result is read by exactly ONE later op is spliced into that consumer as long as
// the merged expression reads at most two distinct buffers, so `add(sin(a), cos(b))` or a binary op
// followed by a unary chain becomes one kernel and one HBM round trip.  The same rules as for unary
// fusing apply to what is skipped: checkpointed buffers, buffers of another length or dtype, buffers a
// grad function needs.  Intermediates that are fused away are never written (they keep the zeros of
// their allocation), exactly like the middle buffers of a fused unary chain (src/op_hint.rs:172-196).
// Call it before optimize_mem_graph (aliasing makes ids share storage).
extern "C" int32_t cbm_elementwise_fusing(cbm_device *d)
{
    CB_CHECK_ARG(d, "null device");
    if (!d->has(CBM_GRAPH)) return fail(CB_ERR_MISSING_CACHE_TRACES, "MissingCacheTraces: no Graph module");
    if (!d->has(CBM_LAZY)) return fail(CB_ERR_UNSUPPORTED, "UnaryFusingUnsupported: only Lazy records operations");
    auto reads = [](const Op &o, uint64_t id) {
        if (o.kind == OpKind::NoOp) return false;
        if (o.kind == OpKind::UnaryGrad) return std::find(o.arg_ids.begin(), o.arg_ids.end(), id) != o.arg_ids.end();
        return std::find(o.in.begin(), o.in.end(), id) != o.in.end();
    };
    auto writes = [](const Op &o, uint64_t id) {
        if (o.kind == OpKind::NoOp) return false;
        if (o.kind == OpKind::UnaryGrad) return o.arg_ids.size() > 1 && o.arg_ids[1] == id;
        return o.out == id;
    };
    auto fusable = [](const Op &o) { return (o.kind == OpKind::Apply || o.kind == OpKind::Binary || o.kind == OpKind::Apply2) && !o.tree.empty(); };
    std::vector<char> changed(d->ops.size(), 0);
    for (size_t c = 0; c < d->ops.size(); c++) {
        bool again = true;
        while (again && fusable(d->ops[c])) {
            again = false;
            Op &oc = d->ops[c];
            for (size_t slot = 0; slot < oc.in.size() && !again; slot++) {
                const uint64_t B = oc.in[slot];
                auto po = d->op_of_id.find(B);
                if (po == d->op_of_id.end() || po->second >= c) continue;
                const size_t p = po->second;
                const Op &op = d->ops[p];
                if (!fusable(op) || op.out != B || op.dtype != oc.dtype) continue;
                const Entry *eb = d->resolve(B);
                auto gi = d->buf_id_to_idx.find(B);
                if (gi == d->buf_id_to_idx.end() || d->graph.node(gi->second).skip) continue;  // checkpoint()
                auto go = d->buf_id_to_idx.find(oc.out);
                if (go == d->buf_id_to_idx.end() || d->graph.node(go->second).len != d->graph.node(gi->second).len) continue;
                (void)eb;
                // exactly one reader, nobody else writes it, no grad function needs it
                size_t n_readers = 0, n_writers = 0;
                for (const Op &o : d->ops) {
                    n_readers += reads(o, B) ? 1 : 0;
                    n_writers += writes(o, B) ? 1 : 0;
                }
                bool on_tape = false;
                for (const GradOp &g : d->tape) on_tape = on_tape || g.buf_id == B || g.out_id == B;
                if (n_readers != 1 || n_writers != 1 || on_tape) continue;
                // moving the producer down to the consumer must not cross a write to one of its inputs
                bool hazard = false;
                for (size_t q = p + 1; q < c && !hazard; q++)
                    for (uint64_t pin : op.in) hazard = hazard || writes(d->ops[q], pin);
                if (hazard) continue;
                // merged input list: the consumer's other input (if any) first, then the producer's
                std::vector<uint64_t> merged;
                for (size_t k = 0; k < oc.in.size(); k++)
                    if (k != slot) merged.push_back(oc.in[k]);
                for (uint64_t pin : op.in)
                    if (std::find(merged.begin(), merged.end(), pin) == merged.end()) merged.push_back(pin);
                if (merged.size() > 2 || op.tree.size() + oc.tree.size() > (size_t)kMaxNodes) continue;
                auto marker_of = [&merged](uint64_t id) {
                    return (int32_t)(std::find(merged.begin(), merged.end(), id) - merged.begin()) == 0 ? CB_OP_X : CB_OP_Y;
                };
                const int32_t inner_map[2] = {marker_of(op.in[0]), op.in.size() > 1 ? marker_of(op.in[1]) : CB_OP_X};
                const int32_t other_to = oc.in.size() > 1 ? marker_of(oc.in[1 - slot]) : CB_OP_X;
                std::vector<cb_node> tree = substitute(oc.tree, slot == 0 ? CB_OP_X : CB_OP_Y, other_to, op.tree, inner_map);
                oc.tree = std::move(tree);
                oc.in = merged;
                d->ops[p] = Op();  // Operation::no_op()
                d->op_of_id.erase(B);
                changed[c] = 1;
                again = true;
            }
        }
    }
    for (size_t c = 0; c < d->ops.size(); c++) {
        if (!changed[c] || d->ops[c].kind == OpKind::NoOp) continue;  // spliced further down itself
        Op &o = d->ops[c];
        const cb_node *progs[1] = {o.tree.data()};
        const int32_t counts[1] = {(int32_t)o.tree.size()};
        if (o.in.size() == 1) {
            CB_TRY(cb_expr_compile(d->raw, o.dtype, CB_KERNEL_APPLY, progs, counts, 1, &o.expr));
            o.kind = OpKind::Apply;
            o.arg_ids = {o.out, o.in[0]};
        } else {
            CB_TRY(cb_expr_compile(d->raw, o.dtype, CB_KERNEL_BINARY, progs, counts, 1, &o.expr));
            o.kind = OpKind::Apply2;
            o.arg_ids = {o.in[0], o.in[1], o.out};
        }
        o.unary_hint = false;
        o.hint.clear();
        o.fused = true;
    }
    d->invalidate_replay();
    return CB_OK;
}

// ===================================================================== Cached: cursor
extern "C" int32_t cbm_cursor(cbm_device *d, uint64_t *cursor)
{
    CB_CHECK_ARG(d && cursor, "null argument");
    *cursor = d->cursor;
    return CB_OK;
}

extern "C" int32_t cbm_set_cursor(cbm_device *d, uint64_t cursor)
{
    CB_CHECK_ARG(d, "null device");
    d->cursor = cursor;
    return CB_OK;
}

// ===================================================================== Autograd
extern "C" int32_t cbm_grad(cbm_device *d, cbm_buf b, cbm_buf *grad)
{
    CB_CHECK_ARG(d && grad, "null argument");
    if (!d->has(CBM_AUTOGRAD)) return fail(CB_ERR_STATE, "Autograd<> is not available.");  // impl_autograd.rs:9
    GET_HANDLE(h, d, b);
    const uint64_t id = h->id;
    Entry *g = nullptr;
    CB_TRY(grad_entry(d, id, h->len, h->dtype, &g));
    auto it = d->grad_handle.find(id);
    if (it != d->grad_handle.end() && d->handle(it->second)) {
        *grad = it->second;
        return CB_OK;
    }
    Handle gh;
    gh.id = g->ptr;
    gh.ptr = g->ptr;
    gh.len = g->len;
    gh.dtype = g->dtype;
    gh.owned = false;  // owned by the gradient pool
    gh.global_len = h->global_len;
    gh.shard_begin = h->shard_begin;
    d->buffers[gh.id] = *g;
    *grad = d->new_handle(gh);
    d->grad_handle[id] = *grad;
    return CB_OK;
}

static int32_t backward_impl(cbm_device *d, cbm_buf out, const void *seed, size_t seed_len)
{
    if (!d->has(CBM_AUTOGRAD)) return CB_OK;  // "should never be None" — without a tape nothing happens
    GET_HANDLE(ho, d, out);
    Entry *og = nullptr;
    CB_TRY(grad_entry(d, ho->id, ho->len, ho->dtype, &og));
    // The seed of ones can be folded into the first grad kernel of the replay (CB_GRAD_SEED_ONES: the kernel computes
    // with 1 and WRITES out.grad instead of reading it) when that kernel is the one consuming out.grad and it will
    // really run; otherwise out.grad is filled by its own pass, like `vec![T::one(); len]` + write in the reference.
    bool seed_in_kernel = false;
    if (!seed && !d->tape.empty()) {
        const GradOp &last = d->tape.back();
        const Entry *buf = d->resolve(last.buf_id);
        const Entry *o = d->resolve(last.out_id);
        auto rg = d->requires_grad.find(last.buf_id);
        seed_in_kernel = last.out_id == ho->id && last.buf_id != ho->id && buf && o && buf->len == og->len &&
                         rg != d->requires_grad.end() && rg->second;
    }
    if (seed) {
        if (seed_len != og->len) return fail(CB_ERR_SHAPE, "seed of %zu elements for a buffer of %zu", seed_len, og->len);
        CB_TRY(cb_h2d(d->raw, og->ptr, seed, seed_len * dtype_size(og->dtype)));  // tape.rs:53-64
    } else if (!seed_in_kernel) {
        CB_TRY(cb_fill(d->raw, og->dtype, og->ptr, og->len, 1.0, 1));  // vec![T::one(); len], on the device
    }
    // device.eagerly(|| tape.backward(..)): grad ops run now even under Lazy (tape.rs:78-82)
    const bool lazy_was_enabled = d->lazy_enabled;
    const bool is_lazy_enabled = d->recording();
    d->lazy_enabled = false;
    int32_t rc = CB_OK;
    for (auto it = d->tape.rbegin(); it != d->tape.rend() && rc == CB_OK; ++it) {
        const GradOp &g = *it;
        const Entry *buf = d->resolve(g.buf_id);
        const Entry *o = d->resolve(g.out_id);
        if (!buf || !o) {
            rc = fail(CB_ERR_INVALID_LAZY_BUF, "InvalidLazyBuf: a buffer of a grad function is not alive");
            break;
        }
        auto rg = d->requires_grad.find(g.buf_id);
        if (rg == d->requires_grad.end() || !rg->second) continue;  // if !buf.requires_grad() { return Ok(()) }
        Entry *bg = nullptr, *outg = nullptr;
        rc = grad_entry(d, g.buf_id, buf->len, buf->dtype, &bg);
        if (rc == CB_OK) rc = grad_entry(d, g.out_id, o->len, o->dtype, &outg);
        // one kernel per grad function; a chain entry stands for the K grad functions of a fused chain
        const uint32_t flags = (seed_in_kernel && it == d->tape.rbegin()) ? CB_GRAD_SEED_ONES : 0u;
        if (rc == CB_OK) rc = cb_unary_grad_ex(d->raw, g.grad_expr, buf->ptr, bg->ptr, outg->ptr, buf->len, flags);
    }
    d->lazy_enabled = lazy_was_enabled;
    if (!is_lazy_enabled) d->tape.clear();  // tape.rs:48-50
    return rc;
}

extern "C" int32_t cbm_backward(cbm_device *d, cbm_buf out)
{
    CB_CHECK_ARG(d, "null device");
    return backward_impl(d, out, nullptr, 0);
}

extern "C" int32_t cbm_backward_with(cbm_device *d, cbm_buf out, const void *seed, size_t len)
{
    CB_CHECK_ARG(d && seed, "null argument");
    return backward_impl(d, out, seed, len);
}

extern "C" int32_t cbm_zero_grad(cbm_device *d)
{
    CB_CHECK_ARG(d, "null device");
    for (auto &kv : d->grads) {
        auto rg = d->requires_grad.find(kv.first);
        if (rg != d->requires_grad.end() && !rg->second) continue;  // gradients.rs:33-38
        CB_TRY(cb_clear(d->raw, kv.second.dtype, kv.second.ptr, kv.second.len));
    }
    return CB_OK;
}

extern "C" int32_t cbm_set_grad_enabled(cbm_device *d, int32_t enabled)
{
    CB_CHECK_ARG(d, "null device");
    d->grad_enabled = enabled != 0;
    return CB_OK;
}
