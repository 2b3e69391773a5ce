/* AddressSanitizer / UBSan driver for the C oracle: a tiny table that rehashes many times under random
 * requests holding several counters each, deletes and re-registrations.  Test infrastructure. */
#include "limitador_oracle.h"
#include <stdio.h>
#include <stdlib.h>
static unsigned long long s = 88172645463325252ULL;
static unsigned long long rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main(void) {
    lo_oracle *o = lo_create(16);
    for (unsigned id = 0; id < 12; id++) lo_limit_set(o, id, id / 4, 5 + id, (id % 3 + 1) * 1000000ull, id % 4 != 3);
    unsigned long long now = 1700000000000000ull, lim = 0;
    for (int i = 0; i < 200000; i++) {
        lo_counter c[4];
        unsigned ns = rnd() % 3, m = 1 + rnd() % 4;
        for (unsigned k = 0; k < m; k++) { c[k].limit_id = ns * 4 + k; c[k]._pad = 0; c[k].key_lo = rnd() % 5000; c[k].key_hi = rnd() % 2; }
        uint32_t fl; uint64_t rem[4], ttl[4];
        now += rnd() % 3000;
        int r = lo_check_and_update(o, c, m, 1 + rnd() % 2, (int)(rnd() & 1), now, &fl, rem, ttl);
        if (r < 0) { printf("error %d\n", r); return 1; }
        lim += r;
        if (i % 50000 == 49999) { uint32_t ids[2] = {1, 6}; lo_delete_counters(o, ids, 2); lo_limit_set(o, 1, 0, 6, 2000000ull, 1); lo_limit_set(o, 6, 1, 11, 1000000ull, 1); }
    }
    printf("ok limited=%llu size=%llu\n", lim, (unsigned long long)lo_size(o));
    lo_destroy(o);
    return 0;
}
