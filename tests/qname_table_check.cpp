// QnameTable (raer_host.h) against std::unordered_map under the access pattern of the parser: find, then insert if
// absent or erase if present, entries dying as the "position" advances, rebuilds dropping the dead.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <unordered_map>
#include "raer_host.h"

int main() {
    raer::QnameTable T;
    struct E { int64_t ordinal; int tid, end; bool keep_a; };
    std::unordered_map<std::string, E> M;
    unsigned rs = 7;
    auto rnd = [&]() { rs = rs * 1103515245u + 12345u; return rs >> 8; };
    int pos = 0, tid = 0, bad = 0;
    for (int64_t it = 0; it < 400000; ++it) {
        pos += rnd() % 3;
        if (it % 90000 == 89999) { ++tid; pos = 0; }
        char name[64];
        const int len = snprintf(name, sizeof(name), "%sread_%u", (rnd() % 50 == 0) ? "a_rather_long_prefix_for_the_arena/" : "", rnd() % 6000);
        const int cur_tid = tid, cur_prev = pos;
        auto dead = [cur_tid, cur_prev](const raer::QnameTable::Ent& e) { return e.tid != cur_tid || e.end < cur_prev; };
        const uint64_t h = raer::QnameTable::hash(name, (size_t)len);
        raer::QnameTable::Ent* e = T.find(name, (size_t)len, h);
        auto mi = M.find(name);
        if (e && dead(*e)) { T.erase(e); e = nullptr; }
        if (mi != M.end() && (mi->second.tid != cur_tid || mi->second.end < cur_prev)) { M.erase(mi); mi = M.end(); }
        if ((e != nullptr) != (mi != M.end())) { ++bad; break; }
        if (!e) {
            const int end = pos + 50 + (int)(rnd() % 200);
            const bool ka = rnd() & 1;
            T.insert(name, (size_t)len, h, it, tid, end, ka, dead);
            M[name] = {it, tid, end, ka};
        } else {
            if (e->ordinal != mi->second.ordinal || e->end != mi->second.end || e->keep_a != mi->second.keep_a || e->tid != mi->second.tid) { ++bad; break; }
            T.erase(e);
            M.erase(mi);
        }
    }
    // everything still alive in the map is found with the same payload
    for (auto& kv : M) {
        raer::QnameTable::Ent* e = T.find(kv.first.c_str(), kv.first.size(), raer::QnameTable::hash(kv.first.c_str(), kv.first.size()));
        if (kv.second.tid == tid && kv.second.end >= pos && (!e || e->ordinal != kv.second.ordinal)) ++bad;
    }
    printf("qname table: %d differences, %zu live, %zu slots, arena %zu bytes\n", bad, T.size(), T.tab.size(), T.arena.size());
    return bad ? 1 : 0;
}
