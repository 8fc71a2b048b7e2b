// Stand-in for the reference's src/impl/vamp/planning/nn.hh -- TEST INFRASTRUCTURE ONLY.
// The real header binds the `nigh` k-d tree (an un-vendored dependency, absent from this image) for the
// planners.  planning/plan.hh includes it but Path / PlanningResult / simplify.hh use nothing from it,
// so for the compiled-reference harness an empty header on the include path (before the reference's own)
// is enough.
#pragma once
