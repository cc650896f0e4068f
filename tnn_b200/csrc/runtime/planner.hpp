// Host-side planning for graphs, parameter slabs, pipeline schedules and model partitions -- the
// integer bookkeeping of the reference's graph_builder.hpp:28-112 (Kahn sort over op nodes),
// graph_context.hpp:39-68 (parameters packed into one aligned slab), coordinator.hpp:165-223
// (microbatch issue order) and partitioner/*.hpp (contiguous stage splits), as plain C++ with no
// dependency on the tensor library, so the same code drives the Python front end and any native host.
#pragma once
#include <algorithm>
#include <cstdint>
#include <deque>
#include <limits>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace tnn {
namespace planner {

// ---- graph order ------------------------------------------------------------------------------------------
// Edge e consumes io nodes inputs[e] and produces outputs[e] (ids are arbitrary non-negative integers).
// Returns the edge indices in execution order: Kahn's algorithm with a FIFO ready queue seeded in declaration
// order, so independent edges keep the order they were declared in.  Throws on two producers of one io node and
// on cycles.
inline std::vector<int64_t> topo_order(const std::vector<std::vector<int64_t>>& inputs,
                                       const std::vector<std::vector<int64_t>>& outputs) {
  const int64_t E = (int64_t)inputs.size();
  if ((int64_t)outputs.size() != E) throw std::invalid_argument("topo_order: inputs/outputs length mismatch");
  std::unordered_map<int64_t, int64_t> producer;
  for (int64_t e = 0; e < E; ++e)
    for (int64_t o : outputs[e])
      if (!producer.emplace(o, e).second) throw std::invalid_argument("io node " + std::to_string(o) + " has two producers");
  std::vector<int64_t> indeg((size_t)E, 0);
  std::vector<std::vector<int64_t>> consumers((size_t)E);
  for (int64_t e = 0; e < E; ++e)
    for (int64_t i : inputs[e]) {
      auto it = producer.find(i);
      if (it != producer.end()) {
        ++indeg[(size_t)e];
        consumers[(size_t)it->second].push_back(e);
      }
    }
  std::deque<int64_t> ready;
  for (int64_t e = 0; e < E; ++e)
    if (indeg[(size_t)e] == 0) ready.push_back(e);
  std::vector<int64_t> order;
  order.reserve((size_t)E);
  while (!ready.empty()) {
    const int64_t e = ready.front();
    ready.pop_front();
    order.push_back(e);
    for (int64_t c : consumers[(size_t)e])
      if (--indeg[(size_t)c] == 0) ready.push_back(c);
  }
  if ((int64_t)order.size() != E) throw std::invalid_argument("graph has a cycle");
  return order;
}

// ---- slab layout ------------------------------------------------------------------------------------------
// Offsets (in elements) of tensors of the given sizes packed back to back, each start rounded up to `align`
// elements; the last entry of the result is the total slab length.
inline std::vector<int64_t> plan_slab(const std::vector<int64_t>& sizes, int64_t align) {
  if (align <= 0) throw std::invalid_argument("plan_slab: align must be positive");
  std::vector<int64_t> off;
  off.reserve(sizes.size() + 1);
  int64_t cur = 0;
  for (int64_t n : sizes) {
    if (n < 0) throw std::invalid_argument("plan_slab: negative size");
    off.push_back(cur);
    cur += (n + align - 1) / align * align;
  }
  off.push_back(cur);
  return off;
}

// Gradient buckets for overlapped reduction: parameters are walked in REVERSE registration order (the order
// backward produces their gradients) and cut whenever a bucket reaches `bucket_elems`.  Returns [begin, end)
// element ranges of the slab, last-produced first.
inline std::vector<std::pair<int64_t, int64_t>> plan_buckets(const std::vector<int64_t>& offsets /*plan_slab result*/,
                                                             int64_t bucket_elems) {
  std::vector<std::pair<int64_t, int64_t>> out;
  if (offsets.size() < 2) return out;
  if (bucket_elems <= 0) {
    out.emplace_back(0, offsets.back());
    return out;
  }
  int64_t end = offsets.back();
  for (int64_t i = (int64_t)offsets.size() - 2; i >= 0; --i) {
    if (end - offsets[(size_t)i] >= bucket_elems || i == 0) {
      out.emplace_back(offsets[(size_t)i], end);
      end = offsets[(size_t)i];
    }
  }
  if (end > 0) out.emplace_back(0, end);
  return out;
}

// ---- pipeline schedules -----------------------------------------------------------------------------------
struct Step {
  char op;        // 'F' forward, 'B' backward
  int64_t mb;     // microbatch id
};

// Non-interleaved 1F1B for stage `rank` of `stages`: min(stages-rank-1, M) warm-up forwards, then one forward +
// one backward per remaining microbatch, then the cool-down backwards.  At most stages-rank microbatches of
// activations are alive on a stage.
inline std::vector<Step> schedule_1f1b(int64_t stages, int64_t rank, int64_t microbatches) {
  if (stages <= 0 || rank < 0 || rank >= stages || microbatches <= 0) throw std::invalid_argument("schedule_1f1b: bad arguments");
  const int64_t warm = std::min(stages - rank - 1, microbatches);
  std::vector<Step> s;
  s.reserve((size_t)(2 * microbatches));
  int64_t f = 0, b = 0;
  for (int64_t i = 0; i < warm; ++i) s.push_back({'F', f++});
  for (int64_t i = 0; i < microbatches - warm; ++i) {
    s.push_back({'F', f++});
    s.push_back({'B', b++});
  }
  for (int64_t i = 0; i < warm; ++i) s.push_back({'B', b++});
  return s;
}

// The reference coordinator's order (GPipe with immediate backward): every forward is injected first, each
// backward is issued as soon as its loss is available -- on a stage that is all forwards, then all backwards.
inline std::vector<Step> schedule_gpipe(int64_t microbatches) {
  if (microbatches <= 0) throw std::invalid_argument("schedule_gpipe: bad arguments");
  std::vector<Step> s;
  for (int64_t i = 0; i < microbatches; ++i) s.push_back({'F', i});
  for (int64_t i = 0; i < microbatches; ++i) s.push_back({'B', i});
  return s;
}

// Peak number of microbatch activations alive at once under a schedule.
inline int64_t peak_in_flight(const std::vector<Step>& s) {
  int64_t cur = 0, peak = 0;
  for (const Step& st : s) {
    cur += st.op == 'F' ? 1 : -1;
    peak = std::max(peak, cur);
  }
  return peak;
}

// ---- partitions -------------------------------------------------------------------------------------------
// Contiguous split of `total` items by integer proportions: every part gets at least one item, the last takes
// the remainder (naive_partitioner.hpp:19-72).  Returns (start, length) pairs.
inline std::vector<std::pair<int64_t, int64_t>> split_by_proportions(int64_t total, const std::vector<int64_t>& props,
                                                                    bool at_least_one) {
  if (props.empty()) throw std::invalid_argument("proportions must not be empty");
  int64_t psum = 0;
  for (int64_t p : props) {
    if (p <= 0) throw std::invalid_argument("proportions must be positive");
    psum += p;
  }
  const int64_t n = (int64_t)props.size();
  if (at_least_one && total < n) throw std::invalid_argument("fewer items than partitions");
  std::vector<std::pair<int64_t, int64_t>> parts;
  int64_t cur = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t size = total * props[(size_t)i] / psum;
    int64_t end = i == n - 1 ? total : cur + size;
    if (at_least_one) end = std::min(std::max(end, cur + 1), total);
    else end = std::min(end, total);
    if (at_least_one && end - cur <= 0) throw std::invalid_argument("partition length is zero, check proportions and model size");
    parts.emplace_back(cur, end - cur);
    cur = end;
  }
  return parts;
}

// Exact min-bottleneck contiguous split of `cost` into n non-empty groups; boundary[j] (optional, same length as
// cost) is charged to a group that ENDS after item j when it is not the last item (the activation that has to
// cross to the next stage).  O(n * L^2) dynamic programme (weighted_partitioner.hpp / ftd_partitioner.hpp).
inline std::vector<std::pair<int64_t, int64_t>> min_bottleneck_split(const std::vector<double>& cost, int64_t n,
                                                                    const std::vector<double>& boundary) {
  const int64_t L = (int64_t)cost.size();
  if (L == 0 || n <= 0) throw std::invalid_argument("min_bottleneck_split: empty input");
  if (!boundary.empty() && (int64_t)boundary.size() != L) throw std::invalid_argument("boundary length mismatch");
  n = std::min(n, L);
  std::vector<double> pre((size_t)L + 1, 0.0);
  for (int64_t i = 0; i < L; ++i) pre[(size_t)i + 1] = pre[(size_t)i] + cost[(size_t)i];
  const double INF = std::numeric_limits<double>::infinity();
  std::vector<std::vector<double>> best((size_t)n + 1, std::vector<double>((size_t)L + 1, INF));
  std::vector<std::vector<int64_t>> cut((size_t)n + 1, std::vector<int64_t>((size_t)L + 1, 0));
  best[0][0] = 0.0;
  for (int64_t k = 1; k <= n; ++k)
    for (int64_t j = k; j <= L - (n - k); ++j)
      for (int64_t i = k - 1; i < j; ++i) {
        if (best[(size_t)k - 1][(size_t)i] == INF) continue;
        double seg = pre[(size_t)j] - pre[(size_t)i];
        if (!boundary.empty() && j < L) seg += boundary[(size_t)j - 1];
        const double v = std::max(best[(size_t)k - 1][(size_t)i], seg);
        if (v < best[(size_t)k][(size_t)j]) {
          best[(size_t)k][(size_t)j] = v;
          cut[(size_t)k][(size_t)j] = i;
        }
      }
  std::vector<std::pair<int64_t, int64_t>> parts;
  int64_t j = L;
  for (int64_t k = n; k >= 1; --k) {
    const int64_t i = cut[(size_t)k][(size_t)j];
    parts.emplace_back(i, j - i);
    j = i;
  }
  std::reverse(parts.begin(), parts.end());
  return parts;
}

}  // namespace planner
}  // namespace tnn
