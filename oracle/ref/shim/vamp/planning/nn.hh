// Stand-in for the reference's src/impl/vamp/planning/nn.hh -- TEST INFRASTRUCTURE ONLY.
// The real header binds the `nigh` k-d tree (kavrakilab/nigh@97130999, an un-vendored dependency absent from this
// image) for the planners.  nigh is an EXACT nearest-neighbour structure, so for the compiled-reference harness an
// exact brute-force search with the same interface returns the same neighbour sets: NN<dim>::insert, size and
// nearest(out, query, k, r) = the (at most) k stored nodes closest to the query among those within distance r,
// in ascending order of distance -- with the distance the reference's own Space defines (nn.hh:48-52:
// (FloatVector(a) - FloatVector(b)).l2_norm()).  Ties in distance have measure zero on the sampled inputs.
#pragma once
#include <algorithm>
#include <cstddef>
#include <utility>
#include <vector>

#include <vamp/vector.hh>

namespace vamp::planning
{
    template <std::size_t dim_>
    struct NNFloatArray
    {
        inline static constexpr auto dim = dim_;
        float *v;
    };

    template <std::size_t dimension>
    struct NNNode
    {
        std::size_t index;
        NNFloatArray<dimension> array;

        [[nodiscard]] inline auto as_vector() const noexcept -> FloatVector<dimension>
        {
            return FloatVector<dimension>(array.v);
        }
    };

    template <std::size_t dimension, std::size_t batch = 128>
    struct NN
    {
        std::vector<NNNode<dimension>> nodes;

        void insert(const NNNode<dimension> &n)
        {
            nodes.push_back(n);
        }

        [[nodiscard]] auto size() const -> std::size_t
        {
            return nodes.size();
        }

        static auto distance(const float *a, const float *b) -> float
        {
            const auto diff = FloatVector<dimension>(a) - FloatVector<dimension>(b);
            return diff.l2_norm();
        }

        void nearest(std::vector<std::pair<NNNode<dimension>, float>> &out, const NNFloatArray<dimension> &q, std::size_t k, float r) const
        {
            out.clear();
            for (const auto &n : nodes)
            {
                const float d = distance(n.array.v, q.v);
                if (d <= r)
                {
                    out.emplace_back(n, d);
                }
            }
            std::stable_sort(out.begin(), out.end(), [](const auto &a, const auto &b) { return a.second < b.second; });
            if (out.size() > k)
            {
                out.resize(k);
            }
        }
    };
}  // namespace vamp::planning
