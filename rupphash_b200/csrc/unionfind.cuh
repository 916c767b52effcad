// unionfind.cuh -- lock-free union-find on the device (replaces the sequential union-find of
// scanner.rs:1780-1807).
//
// Invariant: a root is only ever hooked under a SMALLER root (atomicCAS on the root), so
// parent[x] <= x always holds, the root of every tree is the smallest index of its
// component, and the flattened forest is the canonical labelling "smallest member index"
// whatever the interleaving of threads, tiles or ranks.  Path halving writes only ever
// replace parent[x] by another ancestor of x, so racing plain stores are benign.
// All accesses go to L2 (ld/st .cg): L1 is not coherent across SMs and a stale "I am a
// root" line would make the CAS retry loop spin.
#pragma once
#include <stdint.h>

namespace rh {

__device__ __forceinline__ uint32_t uf_find(uint32_t *parent, uint32_t x) {
    uint32_t p = __ldcg(parent + x);
    while (p != x) {
        uint32_t gp = __ldcg(parent + p);
        if (gp != p) __stcg(parent + x, gp);  // path halving
        x = p;
        p = gp;
    }
    return x;
}

__device__ __noinline__ void uf_unite(uint32_t *parent, uint32_t a, uint32_t b) {
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) {
            uint32_t t = a;
            a = b;
            b = t;
        }
        // a > b: hook root a under the smaller root b
        uint32_t old = atomicCAS(parent + a, a, b);
        if (old == a) return;
        // a stopped being a root in the meantime: retry from its new ancestor
    }
}

}  // namespace rh
