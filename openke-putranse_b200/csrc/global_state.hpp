// Process-global state behind the Base.so-compatible surface (shared by capi_host.cpp and
// compat_device.cu).
#pragma once
#include "graph_host.hpp"

namespace pk {

// The reference keeps everything in process-global variables shared by every loader/tester object
// (SURVEY.md section 8(b) "Ownership"); so does its drop-in.
struct Global {
    Graph graph;
    GlibcRand rng{1};
    int64_t seed = 0;
    uint64_t lcg[64] = {0};
    Universe universe;
    bool have_universe = false;
    bool swapped = false;
    int64_t last_head = 0, last_tail = 0, last_valid_head = 0, last_valid_tail = 0;
    uint64_t eval_epoch = 1;   // bumps whenever importTestFiles replaces the evaluation lists (device filter caches key on it)
    uint64_t index_epoch = 1;  // bumps whenever the sampler's id space changes (device caches key on it)
};

Global& G();
const TripleIndex& current_index();
uint64_t index_epoch();
void test_metrics_reset();   // compat_device.cu
void valid_metrics_reset();

}  // namespace pk
