// Stand-in for orlp/pdqsort@b1ef26a5 (un-vendored dependency of the reference, used only to order
// points during the CAPT build and in the pointcloud filters).  TEST INFRASTRUCTURE ONLY.
// CAPT query results are independent of how ties are ordered (the query is exact), so std::sort
// is a faithful replacement for validation purposes.
#pragma once
#include <algorithm>

template <class It, class Cmp>
inline void pdqsort_branchless(It b, It e, Cmp c)
{
    std::sort(b, e, c);
}

template <class It, class Cmp>
inline void pdqsort(It b, It e, Cmp c)
{
    std::sort(b, e, c);
}

template <class It>
inline void pdqsort(It b, It e)
{
    std::sort(b, e);
}
