// FCIT* (reference planning/fcit.hh:60-360) restated over two batch providers -- host C++, no CUDA in this header.
//
// The reference interleaves three things: a lazy best-first search over the complete graph of the samples drawn so
// far, one validate_motion per edge the search commits to (fcit.hh:230-260), and batches of batch_size new valid
// samples, one fkcc each (fcit.hh:322-348).  The search is sequential by nature; what it CONSUMES is not:
//   * Samples::take(n)      the next n valid samples of the Halton stream.  The GPU provider (vmv_planner.cu) generates
//                           and validates whole chunks of the stream on the device (vmv_validate_halton).
//   * Edges::valid(p, c)    validate_motion(state p, state c).  The GPU provider validates, on the first question
//                           about a parent p, the edges from p to EVERY node in one indexed batch and answers from the
//                           cached row afterwards (the search asks about few parents and many children).
// The control flow below follows the reference statement for statement where the order of operations decides the
// result -- including the scratch configurations it reuses between loop iterations (`self` below) -- so that paths,
// costs and iteration counts are the reference's (tests/test_planner.py, tests/test_planner_cpu.py).
#pragma once
#include <algorithm>
#include <cstdint>
#include <limits>
#include <unordered_set>
#include <vector>

namespace vmvfcit
{
    struct Neighbor
    {
        uint32_t index;
        float cost;
    };

    struct Node
    {
        float g = std::numeric_limits<float>::infinity();
        uint32_t seen = 0;    // nodes [0, seen) have been offered to this node as neighbours (reference: sampleIdx)
        size_t cursor = 0;    // next entry of `neighbors` to queue (reference: neighbor_iterator)
        std::vector<Neighbor> neighbors;
        std::unordered_set<uint32_t> invalid;
    };

    struct QueueEdge
    {
        uint32_t index, parent;
        float cost;
    };

    struct Result
    {
        std::vector<uint32_t> path;  // node indices start .. goal (just the start when unsolved)
        float cost = std::numeric_limits<float>::infinity();
        size_t iterations = 0;
        size_t edges_asked = 0;
    };

    // States: float[n][dof] growing at the back; dist(a, b) must be the reference's Configuration::distance.
    template <typename Dist, typename Samples, typename Edges>
    Result solve(std::vector<float> &states, int dof, size_t max_iterations, size_t max_samples, size_t batch_size, bool optimize, Dist &&dist,
                 Samples &&samples, Edges &&edges)
    {
        constexpr uint32_t kNone = std::numeric_limits<uint32_t>::max();
        constexpr uint32_t kStart = 0, kGoal = 1;
        auto state = [&](uint32_t i) { return states.data() + static_cast<size_t>(i) * dof; };
        std::vector<Node> nodes(states.size() / dof);
        std::vector<uint32_t> parents(nodes.size(), kNone);
        nodes[kStart].g = 0.F;
        std::vector<QueueEdge> open;
        Result R;
        size_t iter = 0;
        uint32_t self = kNone;  // whose state the reference's `temp_config_self` holds (it survives loop iterations)
        const auto by_cost = [](const Neighbor &a, const Neighbor &b) { return a.cost < b.cost; };

        while (nodes.size() < max_samples && iter++ < max_iterations)
        {
            {
                Node &S = nodes[kStart];
                // offer the nodes the start has not seen yet (fcit.hh:148-171)
                for (uint32_t nb = S.seen; nb < nodes.size(); ++nb)
                {
                    if (nb == kStart)
                    {
                        continue;
                    }
                    const float d = dist(state(kStart), state(nb));
                    if (d < nodes[nb].g)
                    {
                        S.neighbors.push_back({nb, d + dist(state(kGoal), state(nb))});
                    }
                }
                S.seen = static_cast<uint32_t>(nodes.size());
                std::sort(S.neighbors.begin(), S.neighbors.end(), by_cost);
                S.cursor = 0;
                if (!S.neighbors.empty())
                {
                    open.push_back({S.neighbors[0].index, kStart, S.neighbors[0].cost});
                    S.cursor = 1;
                }

                while (!open.empty())
                {
                    std::sort(open.begin(), open.end(), [](const QueueEdge &a, const QueueEdge &b) { return a.cost > b.cost; });
                    const QueueEdge cur = open.back();
                    open.pop_back();
                    const uint32_t ci = cur.index, cp = cur.parent;
                    float cur_g = nodes[ci].g;

                    // the parent's next neighbour that could still improve something goes on the queue (fcit.hh:201-218)
                    {
                        Node &P = nodes[cp];
                        while (P.cursor != P.neighbors.size())
                        {
                            const Neighbor nb = P.neighbors[P.cursor];
                            ++P.cursor;
                            if (nb.cost < nodes[nb.index].g + dist(state(kGoal), state(nb.index)))
                            {
                                open.push_back({nb.index, cp, nb.cost});
                                break;
                            }
                        }
                    }

                    if (parents[ci] != cp)
                    {
                        self = ci;
                        const float to_goal = dist(state(kGoal), state(ci));
                        if (!(cur.cost <= nodes[kGoal].g))
                        {
                            break;  // nothing left on the queue can beat the solution (fcit.hh:262-265)
                        }
                        if (cur.cost < cur_g + to_goal && nodes[ci].invalid.count(cp) == 0)
                        {
                            ++R.edges_asked;
                            const bool ok = ci == cp || edges(cp, ci);
                            if (ok)
                            {
                                parents[ci] = cp;
                                cur_g = nodes[cp].g + dist(state(cp), state(ci));
                                nodes[ci].g = cur_g;
                            }
                            else
                            {
                                nodes[cp].invalid.insert(ci);
                                nodes[ci].invalid.insert(cp);
                                // the reference marks the entry BEFORE the parent's cursor, wherever that now stands
                                Node &P = nodes[cp];
                                if (P.cursor >= 1)
                                {
                                    P.neighbors[P.cursor - 1].cost = std::numeric_limits<float>::max();
                                }
                                continue;
                            }
                        }
                    }

                    // expansion: every node the current one has not seen becomes its neighbour (fcit.hh:270-310)
                    Node &C = nodes[ci];
                    bool added = false;
                    for (uint32_t nb = C.seen; nb < nodes.size(); ++nb)
                    {
                        if (nb == ci)
                        {
                            continue;
                        }
                        // (`self`, not ci: the reference computes this distance from its scratch configuration)
                        const float d = dist(state(self), state(nb));
                        C.neighbors.push_back({nb, cur_g + d + dist(state(kGoal), state(nb))});
                        added = true;
                    }
                    C.seen = static_cast<uint32_t>(nodes.size());
                    if (added)
                    {
                        std::sort(C.neighbors.begin(), C.neighbors.end(), by_cost);
                        C.cursor = 1;
                        open.push_back({C.neighbors[0].index, ci, C.neighbors[0].cost});
                    }
                }
            }

            if (!optimize && parents[kGoal] != kNone)
            {
                break;
            }

            // a batch of new valid samples (fcit.hh:322-348)
            const size_t room = max_samples - nodes.size();
            const size_t got = samples(std::min(batch_size, room), states);
            nodes.resize(nodes.size() + got);
            parents.resize(nodes.size(), kNone);
            if (got == 0 && batch_size > 0 && room > 0)
            {
                break;  // the sample stream is exhausted
            }
        }

        R.iterations = iter;
        R.cost = nodes[kGoal].g;
        for (uint32_t c = kGoal; parents[c] != kNone; c = parents[c])
        {
            R.path.push_back(c);
        }
        R.path.push_back(kStart);
        std::reverse(R.path.begin(), R.path.end());
        return R;
    }
}  // namespace vmvfcit
