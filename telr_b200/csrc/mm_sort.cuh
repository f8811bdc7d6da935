// mm_sort.cuh — order-exact emulation of klib's radix_sort (ksort.h KRADIX_SORT_INIT) as used by
// minimap2 for anchors, chains and regions: <=64 elements insertion sort (stable); otherwise an
// in-place MSD radix sort on 8-bit digits whose cycle-leader permutation is unstable.  The tie
// order it leaves is observable downstream, so the permutation is reproduced step for step; the
// recursion is unrolled onto an explicit stack and levels whose digit is constant over a range are
// skipped (they are no-ops upstream).
#pragma once
#include "mm_types.cuh"

namespace telr {

struct KeyX { TELR_HD uint64_t operator()(const Anchor &a) const { return a.x; } };
struct KeyId { TELR_HD uint64_t operator()(const uint64_t &a) const { return a; } };

template <class T, class K> TELR_HD void ins_sort(T *a, int n, K key)
{
    for (int i = 1; i < n; ++i)
        if (key(a[i]) < key(a[i - 1])) {
            T tmp = a[i];
            int j;
            for (j = i; j > 0 && key(tmp) < key(a[j - 1]); --j) a[j] = a[j - 1];
            a[j] = tmp;
        }
}

// scratch: int32 words; needs 512 + 3 * (n / 65 + 10)
template <class T, class K> TELR_HDN void rs_sort_emul(T *a, int n, K key, int32_t *scratch)
{
    if (n <= 64) { ins_sort(a, n, key); return; }
    int32_t *bb = scratch, *be = scratch + 256, *stk = scratch + 512;
    int sp = 0;
    stk[0] = 0, stk[1] = n, stk[2] = 56, sp = 1;
    while (sp > 0) {
        --sp;
        int beg = stk[3 * sp], end = stk[3 * sp + 1], s = stk[3 * sp + 2];
        // skip levels on which every key has the same digit
        for (;;) {
            uint64_t k0 = key(a[beg]) >> s & 255;
            bool same = true;
            for (int i = beg + 1; i < end; ++i)
                if ((key(a[i]) >> s & 255) != k0) { same = false; break; }
            if (!same || s == 0) break;
            s -= 8;
        }
        for (int k = 0; k < 256; ++k) bb[k] = be[k] = 0;
        for (int i = beg; i < end; ++i) ++be[key(a[i]) >> s & 255];
        {
            int acc = beg;
            for (int k = 0; k < 256; ++k) { int c = be[k]; bb[k] = acc; acc += c; be[k] = acc; }
        }
        for (int k = 0; k < 256;) {
            if (bb[k] != be[k]) {
                int l = (int)(key(a[bb[k]]) >> s & 255);
                if (l != k) {
                    T tmp = a[bb[k]], swp;
                    do {
                        swp = tmp; tmp = a[bb[l]]; a[bb[l]++] = swp;
                        l = (int)(key(tmp) >> s & 255);
                    } while (l != k);
                    a[bb[k]++] = tmp;
                } else ++bb[k];
            } else ++k;
        }
        if (s) {
            int prev = beg;
            for (int k = 0; k < 256; ++k) {
                int e = be[k], sz = e - prev;
                if (sz > 64) { stk[3 * sp] = prev, stk[3 * sp + 1] = e, stk[3 * sp + 2] = s - 8; ++sp; }
                else if (sz > 1) ins_sort(a + prev, sz, key);
                prev = e;
            }
        }
    }
}

TELR_HD int rs_scratch_words(int n) { return 512 + 3 * (n / 65 + 10); }

}  // namespace telr
