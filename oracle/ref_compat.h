// TEST INFRASTRUCTURE -- not part of the shipped product.
//
// Force-included (-include) when oracle/Makefile compiles the reference's
// own src/render.cpp from where it lies under /root/reference.  The stock file
// does not build with GCC: src/integrator/path_tracing.h:34,192,304 bind
// `auto& light_point = sample_on_light(...)` (a prvalue) to a non-const lvalue
// reference, which MSVC accepts as an extension.  Instead of patching a copy
// of the source we route that one call through a helper that parks the
// returned value in thread-local storage and hands back an lvalue reference
// to it -- the integrator sees the same value, nothing else changes.
#pragma once
#include "scene.h"  // pulls in light.h (the real sample_on_light) before the macro exists

namespace take_oracle_compat {
inline PointAndNormal &hold(const PointAndNormal &p) {
    static thread_local PointAndNormal slot;
    slot = p;
    return slot;
}
}  // namespace take_oracle_compat

// A function-like macro is not re-expanded inside its own replacement list, so
// the inner name below is the reference's inline function from light.h:54-56.
#define sample_on_light(scene_, light_, ref_pos_, rng_) \
    take_oracle_compat::hold(sample_on_light(scene_, light_, ref_pos_, rng_))
