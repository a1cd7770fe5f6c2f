/* A plain C caller of the C ABI (include/fss.h): what the reference-side shim of INTEGRATION.md does, without any
 * Python in between.  Built by tests/test_abi.py with gcc and linked against libfss_b200.so; run on the GPU box by
 * tests/test_gpu_parity.py.  Prints a digest that the test compares with the Python binding's. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fss.h"
#include "fss_rng.h"

#define CHECK(call)                                                                 \
    do {                                                                            \
        int rc_ = (call);                                                           \
        if(rc_ != FSS_OK) {                                                         \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, fss_last_error());        \
            return rc_ == FSS_ERR_NO_DEVICE ? 3 : 1;                                \
        }                                                                           \
    } while(0)

int main(int argc, char** argv) {
    const int W = 264, H = 208, ticks = argc > 1 ? atoi(argv[1]) : 8;
    fss_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    cfg.width = W;
    cfg.height = H;
    cfg.seed = 0xF55;
    cfg.flags = FSS_PARTICLES_EMIT;
    cfg.condense_material = 3;
    fss_world* w = NULL;
    CHECK(fss_create(&cfg, &w));

    /* AIR, SOLID, a sand, water, steam: Materials.cpp:5-46 in miniature */
    fss_material m[5];
    memset(m, 0, sizeof m);
    m[0].physics_type = FSS_PHYS_AIR;   m[0].conduction_self = m[0].conduction_other = 0.8f;
    m[1].physics_type = FSS_PHYS_SOLID; m[1].density = 1.0f; m[1].conduction_self = m[1].conduction_other = 1.0f; m[1].create_base = 0x808080;
    m[2].physics_type = FSS_PHYS_SAND;  m[2].density = 10.0f; m[2].iterations = 2; m[2].slipperyness = 20;
    m[2].conduction_self = m[2].conduction_other = 1.0f;
    m[2].create_kind = FSS_CREATE_RAND; m[2].create_base = (220 << 16) + (155 << 8) + 100; m[2].create_rand_mod = 30; m[2].create_rand_shift = 8;
    m[2].create_rand_site = FSS_SITE_TILES + 16;
    m[3].physics_type = FSS_PHYS_SOUP;  m[3].density = 1.5f; m[3].iterations = 6; m[3].conduction_self = m[3].conduction_other = 1.0f;
    m[3].create_base = 0x00B69F; m[3].create_temperature = -1023;
    m[4].physics_type = FSS_PHYS_GAS;   m[4].density = -1.0f; m[4].iterations = 1; m[4].flags = FSS_MAT_IS_STEAM;
    m[4].conduction_self = m[4].conduction_other = 1.0f; m[4].create_base = 0x666666;
    CHECK(fss_set_materials(w, m, 5));
    CHECK(fss_set_tick_zone(w, 16, 16, W - 32, H - 48));

    fss_fill_spec spec;
    memset(&spec, 0, sizeof spec);
    spec.seed = 7; spec.border = 16; spec.border_material = 1; spec.n_entries = 4;
    spec.material[0] = 2; spec.cumulative_percent[0] = 25;
    spec.material[1] = 3; spec.cumulative_percent[1] = 45;
    spec.material[2] = 1; spec.cumulative_percent[2] = 55;
    spec.material[3] = 4; spec.cumulative_percent[3] = 60;
    CHECK(fss_fill_synthetic(w, &spec));

    /* a host edit through the mirror: a block of sand */
    fss_cell* patch = (fss_cell*)calloc(20 * 10, sizeof(fss_cell));
    CHECK(fss_download_rect(w, 100, 60, 20, 10, patch, 20));
    for(int i = 0; i < 200; i++) { patch[i].material = 2; patch[i].fluid_amount = 2.0f; patch[i].fluid_amount_diff = 0; patch[i].moved = 0; }
    CHECK(fss_upload_rect(w, 100, 60, 20, 10, patch, 20));

    for(int t = 0; t < ticks; t++) {
        CHECK(fss_tick(w, (uint32_t)t));
        if(t % 4 == 2) CHECK(fss_tick_temperature(w)); /* Game.cpp:2759 */
    }
    uint64_t digest = 0, counts[5];
    size_t n = 0, total = 0;
    fss_particle_spawn* sp = (fss_particle_spawn*)malloc(sizeof(fss_particle_spawn) * 65536);
    CHECK(fss_hash(w, &digest));
    CHECK(fss_histogram(w, counts, 5));
    CHECK(fss_drain_particles(w, sp, 65536, &n, &total));
    fss_stats st;
    CHECK(fss_get_stats(w, &st));
    printf("digest %016llx particles %zu counts %llu %llu %llu %llu %llu launches %llu\n", (unsigned long long)digest, total,
           (unsigned long long)counts[0], (unsigned long long)counts[1], (unsigned long long)counts[2], (unsigned long long)counts[3],
           (unsigned long long)counts[4], (unsigned long long)st.kernel_launches);
    free(patch);
    free(sp);
    CHECK(fss_destroy(w));
    return 0;
}
