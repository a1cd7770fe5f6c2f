/*
 * fss_oracle.h -- plain-C restatement of World::tick / World::tickTemperature.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may build, load or call this.  The product (fallingsandsurvival_b200/) never does.
 *
 * Parity of THIS file is pinned against the reference's own lines compiled headless
 * (oracle/_ref, built by oracle/build_ref.py): tests/test_oracle_vs_reference.py runs both on the
 * same seeded worlds and requires bit-identical grids.  The reference ships no tests or golden
 * vectors for this path, so the goldens in tests/golden/ were minted from
 * oracle/_ref by tests/golden/make_golden.py.
 */
#ifndef FSS_ORACLE_H
#define FSS_ORACLE_H

#include "../include/fss.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fsso_world fsso_world;

fsso_world* fsso_create(int width, int height, uint32_t seed, uint32_t flags, int condense_material);
void fsso_destroy(fsso_world* w);
int fsso_set_materials(fsso_world* w, const fss_material* table, int n);
int fsso_set_texture(fsso_world* w, int index, const uint32_t* argb, int tw, int th);
int fsso_set_tick_zone(fsso_world* w, int x, int y, int zw, int zh);
fss_cell* fsso_cells(fsso_world* w); /* row-major width*height, owned by the world */
float* fsso_flow_x(fsso_world* w);   /* World::flowX / flowY, row-major width*height: the caller zeroes them when it consumes them */
float* fsso_flow_y(fsso_world* w);
uint8_t* fsso_dirty(fsso_world* w);  /* World::dirty, row-major width*height */
int fsso_set_material_looks(fsso_world* w, const uint8_t* alpha, const uint32_t* emit_color, int n);
/* the `dirty` job of Game::tick (F/Game.cpp:2579-2650) + the memset of :2751; pixels[k]: width*height*4 bytes each */
void fsso_render_dirty(fsso_world* w, uint8_t* const pixels[4], uint64_t* moving, int n_materials, fss_render_info* info);

void fsso_tick(fsso_world* w, uint32_t tick_ct);
/* one (iter, tk) sub-step restricted to micro-chunks whose first row lies in [cy_from, cy_to) */
void fsso_substep_rows(fsso_world* w, uint32_t tick_ct, int iter, int tk, int cy_from, int cy_to);
void fsso_tick_temperature(fsso_world* w);
/* temperature update restricted to rows [y_from, y_to) (reads one row beyond) */
void fsso_tick_temperature_rows(fsso_world* w, int y_from, int y_to);

size_t fsso_num_particles(fsso_world* w);
size_t fsso_drain_particles(fsso_world* w, fss_particle_spawn* out, size_t cap); /* sorted like fss_drain_particles */
/* World::particles and World::tickParticles (F/world.cpp:2171-2311); the tick first adopts the spawn records */
void fsso_particles_add(fsso_world* w, const fss_particle* p, size_t n);
fss_particle* fsso_particles(fsso_world* w, size_t* n);
void fsso_tick_particles(fsso_world* w);

uint64_t fsso_hash_rows(const fss_cell* cells, int width, int x0, int y0, int w, int h, int global_y0);
void fsso_histogram(const fss_cell* cells, size_t n, uint64_t* counts, int n_materials);

#ifdef __cplusplus
}
#endif
#endif
