// CPU check of the descent records' leaf-id encoding (csrc/common.cuh: csp_encode_leaf_runs) against the decode rule the descent
// kernels apply (csrc/locate_eval.cu, descend<2>): for every child pattern of a degree-2 node -- internal / real leaf / fake
// leaf, fictive first or second dimension, arbitrary leaf counts under the internal children -- every reachable leaf child must
// decode to its preorder leaf id, or the node must fall back to the dleaf table (G = ~first descent leaf id).
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"

static int decode(unsigned mask, unsigned idx, int G, int H, bool* table) {  // mirrors descend<2>'s leaf branch
    const unsigned below = (1u << idx) - 1u;
    const unsigned cnt = __builtin_popcount(~mask & below & 0xfu);
    *table = G < 0;
    if (G >= 0) {
        const unsigned lead = mask & ~(mask + 1u);
        return ((mask & below & ~lead) ? H : G) + (int)cnt;
    }
    return ~G + (int)cnt;  // descent leaf id: looked up in dleaf
}

int main() {
    long checked = 0, direct = 0, table_nodes = 0;
    for (unsigned mask = 0; mask < 15; ++mask)           // internal-children mask (15: no leaf child)
        for (unsigned fict = 0; fict < 4; ++fict)          // bit 0: first split dimension fictive, bit 1: second
            for (int s0 = 1; s0 <= 3; ++s0)
                for (int s1 = 1; s1 <= 3; ++s1)
                    for (int first = 0; first <= 5; first += 5) {
                        const unsigned f0 = fict & 1u, f1 = (fict >> 1) & 1u;
                        int L[4] = {-1, -1, -1, -1}, dl[4] = {-1, -1, -1, -1};
                        int next = first, gdl = 7, g = gdl, si = 0;
                        const int sub[2] = {s0, s1};
                        bool valid = true;
                        for (int k = 0; k < 4; ++k) {
                            const bool fake = (f0 && (k & 1)) || (f1 && (k & 2));
                            if ((mask >> k) & 1u) {
                                if (fake) valid = false;  // a child displaced along a fictive dimension is never split here
                                next += sub[si++ & 1];
                            } else {
                                dl[k] = g++;
                                if (!fake) L[k] = next++;
                            }
                        }
                        if (!valid) continue;
                        int G, H;
                        csp_encode_leaf_runs(L, mask, f0, f1, gdl, &G, &H);
                        bool any_table = false;
                        for (unsigned k = 0; k < 4; ++k) {
                            if ((mask >> k) & 1u) continue;
                            if ((f0 && (k & 1)) || (f1 && (k & 2))) continue;  // unreachable
                            bool table;
                            const int id = decode(mask, k, G, H, &table);
                            any_table |= table;
                            const int want = table ? dl[k] : L[k];
                            if (id != want) {
                                printf("MISMATCH mask=%u fict=%u k=%u: got %d want %d (G=%d H=%d)\n", mask, fict, k, id, want, G, H);
                                return 1;
                            }
                            ++checked;
                        }
                        any_table ? ++table_nodes : ++direct;
                        if (fict == 0 && any_table) { printf("node without fake leaves fell back to the table: mask=%u\n", mask); return 1; }
                    }
    printf("ok %ld %ld %ld\n", checked, direct, table_nodes);
    return 0;
}
