// reorder.h -- locality ordering of the element graph (host code).
//
// Greedy graph growing ("METIS-style" initial partitioning without the
// refinement pass): breadth-first patches of `patch` elements are grown one
// after another, each new patch seeded from the frontier of the previous ones,
// so that the threads of one CTA own a compact piece of the mesh and most of
// their neighbour gathers stay inside the CTA's own cache lines.
#pragma once
#include <deque>
#include <vector>

namespace pb {

// nabr: [3][ne] reference neighbour table (1-based, 0 boundary, <0 -(river));
// lr:   [2][nr] left / right bank elements (1-based).
// perm_out[internal] = reference element index (0-based).
inline void patch_order(int ne, int nr, const int *nabr, const int *lr, int patch, int *perm_out)
{
    auto neighbours = [&](int e, int out[3]) {
        int k = 0;
        for (int j = 0; j < 3; j++) {
            const int n = nabr[(size_t)j * ne + e];
            if (n > 0) out[k++] = n - 1;
            else if (n < 0) {
                const int r = -n - 1;
                if (r < nr) {
                    const int l = lr[r] - 1, rt = lr[nr + r] - 1;
                    const int o = (l == e) ? rt : l;
                    if (o >= 0 && o < ne && o != e) out[k++] = o;
                }
            }
        }
        return k;
    };
    std::vector<char> seen(ne, 0);
    std::deque<int> seeds;          // frontier elements left over by finished patches
    int next_unseen = 0, filled = 0;
    std::vector<int> q;
    q.reserve(patch * 2);
    while (filled < ne) {
        int seed = -1;
        while (!seeds.empty()) {
            const int s = seeds.front();
            seeds.pop_front();
            if (!seen[s]) { seed = s; break; }
        }
        if (seed < 0) {
            while (seen[next_unseen]) next_unseen++;
            seed = next_unseen;
        }
        q.clear();
        q.push_back(seed);
        seen[seed] = 1;
        size_t head = 0;
        int taken = 0;
        while (head < q.size() && taken < patch) {
            const int e = q[head++];
            perm_out[filled++] = e;
            taken++;
            int nb[3];
            const int k = neighbours(e, nb);
            for (int t = 0; t < k; t++)
                if (!seen[nb[t]]) { seen[nb[t]] = 1; q.push_back(nb[t]); }
        }
        // elements discovered but not taken go back to the pool as future seeds
        for (size_t t = head; t < q.size(); t++) { seen[q[t]] = 0; seeds.push_back(q[t]); }
    }
}

}  // namespace pb
