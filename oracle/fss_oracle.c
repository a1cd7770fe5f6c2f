/*
 * fss_oracle.c -- plain-C restatement of the reference's per-tick cellular automaton.
 *
 * TEST INFRASTRUCTURE ONLY (see fss_oracle.h).  Straight-line, array-of-structs, one cell at a
 * time, in the reference's scan order; no attempt at speed beyond an OpenMP loop over the
 * micro-chunks of one phase (which the schedule makes independent).
 *
 * Every function cites the reference lines it follows; F/ = /root/reference/FallingSandSurvival/.
 * Schedule = F/world.cpp:1108-1141 with CHUNK_W x CHUNK_H = FSS_CHUNK_W x FSS_CHUNK_H (4 x 16)
 * and rand() replaced by fss_rand() (include/fss_rng.h), exactly what oracle/_ref compiles.
 */
#include "fss_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../include/fss_rng.h"

#define CW FSS_CHUNK_W
#define CH FSS_CHUNK_H

/* F/world.hpp:76-87 */
#define FLUID_MaxValue 0.5f
#define FLUID_MinValue 0.0005f
#define FLUID_MaxCompression 0.1f
#define FLUID_MinFlow 0.05f
#define FLUID_MaxFlow 8.0f
#define FLUID_FlowSpeed 1.0f

typedef struct {
    uint32_t* px;
    int w, h;
} texture_t;

struct fsso_world {
    int width, height;
    int zx, zy, zw, zh;
    uint32_t seed;
    uint32_t flags;
    int condense_material;
    fss_cell* cells;
    float *flow_x, *flow_y; /* World::flowX / flowY (F/world.hpp:99-100) */
    uint8_t* dirty;         /* World::dirty (F/world.hpp:139): set wherever the reference sets it, cleared by the render pass */
    float *prev_flow_x, *prev_flow_y; /* World::prevFlowX / prevFlowY (F/world.hpp:101-102) */
    uint8_t alpha[FSS_MAX_MATERIALS]; /* Material::alpha, Material::emitColor (F/Material.hpp:34, :38) */
    uint32_t emit_color[FSS_MAX_MATERIALS];
    int32_t* new_temps;
    fss_material mat[FSS_MAX_MATERIALS];
    int max_stability[FSS_MAX_MATERIALS];
    int n_mat;
    int fire_id;
    texture_t tex[FSS_MAX_TEXTURES];
    fss_particle_spawn* parts;
    size_t n_parts, cap_parts;
    fss_particle* plist; /* World::particles */
    size_t n_plist, cap_plist;
};

/* per-visit RNG state: one key per cell visit, one counter per site (fss_rng.h) */
typedef struct {
    uint32_t cellkey;
    int n;
    uint32_t site[8];
    uint32_t count[8];
} rng_t;

static uint32_t rnd(rng_t* r, uint32_t site) {
    uint32_t k = 0;
    int i;
    for(i = 0; i < r->n; i++) {
        if(r->site[i] == site) {
            k = ++r->count[i];
            break;
        }
    }
    if(i == r->n && r->n < 8) {
        r->site[r->n] = site;
        r->count[r->n] = 0;
        r->n++;
    }
    return fss_rand(r->cellkey, site, k);
}

/* one micro-chunk's view of the sub-step */
typedef struct {
    fsso_world* w;
    int cx, cy;
    uint32_t tick_ct;
    int iter;
    uint32_t tkey;
    uint8_t vis[CH][CW]; /* tickVisited restricted to the chunk: flags set outside are never read
                            (the other buffer is cleared before reuse, F/world.cpp:1125-1131, :1998) */
} chunk_t;

static inline fss_cell* at(fsso_world* w, int x, int y) { return &w->cells[(size_t)y * w->width + x]; }
static inline int phys(fsso_world* w, const fss_cell* c) { return w->mat[c->material].physics_type; }
/* World::getTile, F/world.cpp:1053-1056: out of bounds reads Tiles::TEST_SOLID */
static inline int phys_get(fsso_world* w, int x, int y) {
    if(x < 0 || x >= w->width || y < 0 || y >= w->height) return FSS_PHYS_SOLID;
    return phys(w, at(w, x, y));
}
static inline void set_dirty(fsso_world* w, int x, int y) { w->dirty[(size_t)y * w->width + x] = 1; }
static inline void mark(chunk_t* c, int x, int y) {
    int dx = x - c->cx, dy = y - c->cy;
    if(dx >= 0 && dx < CW && dy >= 0 && dy < CH) c->vis[dy][dx] = 1;
}

/* Tiles::NOTHING, F/Tiles.cpp:7; fluidAmount keeps the constructor default 2.0f (F/MaterialInstance.hpp:21) */
static fss_cell nothing(void) {
    fss_cell c;
    memset(&c, 0, sizeof c);
    c.fluid_amount = 2.0f;
    return c;
}

/* Tiles::create(mat, x, y), F/Tiles.cpp:190-268, from the table's recipe */
static fss_cell create_cell(fsso_world* w, int m, int x, int y, rng_t* r) {
    const fss_material* d = &w->mat[m];
    fss_cell c = nothing();
    c.material = (uint16_t)m;
    c.temperature = d->create_temperature;
    if(d->create_kind == FSS_CREATE_RAND) {
        c.color = d->create_base + ((rnd(r, d->create_rand_site) % d->create_rand_mod) << d->create_rand_shift);
    } else if(d->create_kind == FSS_CREATE_TEXTURE && d->create_texture >= 0 && w->tex[d->create_texture].px) {
        texture_t* t = &w->tex[d->create_texture];
        c.color = t->px[(y % t->h) * t->w + (x % t->w)];
    } else {
        c.color = d->create_base;
    }
    return c;
}

/* `MaterialInstance(tile.mat, tile.color, tile.temperature)` + fluidAmount = 0  (F/world.cpp:1399-1400 etc.) */
static fss_cell liquid_seed(const fss_cell* tile) {
    fss_cell c = nothing();
    c.material = tile->material;
    c.color = tile->color;
    c.temperature = tile->temperature;
    c.fluid_amount = 0.0f;
    return c;
}

static void emit(chunk_t* c, int x, int y, uint32_t site, uint32_t cellkey, const fss_cell* cell) {
    fsso_world* w = c->w;
    if(!(w->flags & FSS_PARTICLES_EMIT)) return;
#pragma omp critical(fsso_parts)
    {
        if(w->n_parts == w->cap_parts) {
            w->cap_parts = w->cap_parts ? w->cap_parts * 2 : 1024;
            w->parts = (fss_particle_spawn*)realloc(w->parts, w->cap_parts * sizeof(fss_particle_spawn));
        }
        fss_particle_spawn* p = &w->parts[w->n_parts++];
        p->x = (uint32_t)x;
        p->y = (uint32_t)y;
        p->site = site;
        p->cell_key = cellkey;
        p->tick_iter = c->tick_ct * 8u + (uint32_t)c->iter;
        p->cell = *cell;
    }
}

/* F/world.cpp:1075-1088 */
static float vertical_flow(float remaining, float dest) {
    float sum = remaining + dest;
    float value;
    if(sum <= FLUID_MaxValue) {
        value = FLUID_MaxValue;
    } else if(sum < 2 * FLUID_MaxValue + FLUID_MaxCompression) {
        value = (FLUID_MaxValue * FLUID_MaxValue + sum * FLUID_MaxCompression) / (FLUID_MaxValue + FLUID_MaxCompression);
    } else {
        value = (sum + FLUID_MaxCompression) / 2.0f;
    }
    return value;
}

/* the clamp every flow goes through: `flow = std::max(flow, 0.0f); if(flow > std::min(MaxFlow, lim)) flow = ...` */
static float clamp_flow(float flow, float lim) {
    float cap;
    if(flow < 0.0f) flow = 0.0f;
    cap = lim < FLUID_MaxFlow ? lim : FLUID_MaxFlow;
    if(flow > cap) flow = cap;
    return flow;
}

static int can_move_into(fsso_world* w, const fss_cell* dst, const fss_cell* self) {
    int t = phys(w, dst);
    return t == FSS_PHYS_AIR || (t != FSS_PHYS_SOLID && w->mat[dst->material].density < w->mat[self->material].density);
}

/* ---- scan 1, F/world.cpp:1154-1665 ----------------------------------------------------------- */
static void scan1_cell(chunk_t* c, int x, int y) {
    fsso_world* w = c->w;
    int particles = (w->flags & FSS_PARTICLES_EMIT) != 0;
    fss_cell* self = at(w, x, y);
    fss_cell tile;
    int type, dx = x - c->cx, dy = y - c->cy;
    size_t index = (size_t)y * w->width + x;
    rng_t r;

    if(c->vis[dy][dx]) return;                              /* :1161 */
    if(c->iter >= w->mat[self->material].iterations) {      /* :1163-1166 */
        c->vis[dy][dx] = 1;
        return;
    }
    tile = *self;                                           /* :1167 */
    type = phys(w, &tile);
    r.cellkey = fss_rng_cell_key(c->tkey, (uint32_t)x, (uint32_t)y);
    r.n = 0;

    if(w->mat[tile.material].flags & FSS_MAT_IS_FIRE) {     /* :1171-1216 */
        if(rnd(&r, 1172) % 10 == 0) {                       /* recolours the local copy only */
            tile.color = (((255u << 8) + 100 + rnd(&r, 1174) % 50) << 8) + 50;
        }
        if(rnd(&r, 1179) % 10 == 0) emit(c, x, y, 1180, r.cellkey, &tile); /* spark, no grid effect */
        if(rnd(&r, 1191) % 150 == 0) {
            *self = nothing();
            set_dirty(w, x, y);                             /* :1194 */
            c->vis[dy][dx] = 1;
        } else {
            int found = 0, xx, yy;
            for(xx = -2; xx <= 2; xx++) {
                for(yy = -2; yy <= 2; yy++) {
                    fss_cell* n = at(w, x + xx, y + yy);
                    if(phys(w, n) == FSS_PHYS_SOLID) {
                        found = 1;
                        if(rnd(&r, 1202) % 500 == 0) {
                            *n = create_cell(w, w->fire_id, x + xx, y + yy, &r); /* Tiles::createFire() */
                            set_dirty(w, x + xx, y + yy);   /* :1204 */
                            mark(c, x + xx, y + yy);
                        }
                    }
                }
            }
            if(!found && rnd(&r, 1210) % 120 == 0) {
                *self = nothing();
                set_dirty(w, x, y);                         /* :1212 */
                c->vis[dy][dx] = 1;
            }
        }
    }

    if(type == FSS_PHYS_SAND) {                             /* :1218-1337 */
        fss_cell below = *at(w, x, y + 1);
        fss_cell bl, br;
        const fss_material* m = &w->mat[tile.material];
        int can_b, can_l, can_r, i;
        /* :1223-1249 interactions: `interact` is false for every material (F/Materials.cpp:128,:144) */
        if(m->n_reactions > 0) {                            /* :1251-1274 */
            int react = 0;
            for(i = 0; i < m->n_reactions; i++) {
                const fss_reaction* in = &m->reactions[i];
                if((in->type == FSS_REACT_TEMPERATURE_BELOW && tile.temperature < in->data1) ||
                   (in->type == FSS_REACT_TEMPERATURE_ABOVE && tile.temperature > in->data1)) {
                    *self = create_cell(w, in->data2, x, y, &r);
                    self->temperature = tile.temperature;
                    set_dirty(w, x, y);                     /* :1259, :1267 */
                    c->vis[dy][dx] = 1;
                    react = 1;
                }
            }
            if(react) return;
        }
        can_b = can_move_into(w, &below, &tile);            /* :1276 */
        if(!can_b) return;
        bl = *at(w, x - 1, y + 1);
        br = *at(w, x + 1, y + 1);
        can_l = can_move_into(w, &bl, &tile);
        can_r = can_move_into(w, &br, &tile);
        if(!((can_l || can_r) && rnd(&r, 1287) % 20 == 0)) { /* :1287 */
            if(particles && phys(w, &below) == FSS_PHYS_AIR && phys_get(w, x, y + 2) == FSS_PHYS_AIR &&
               phys_get(w, x, y + 3) == FSS_PHYS_AIR && phys_get(w, x, y + 4) == FSS_PHYS_AIR) {
                *self = below;                              /* :1289 setTile(x, y, belowTile), which sets dirty (:1061) */
                set_dirty(w, x, y);
                emit(c, x, y, 1291, r.cellkey, &tile);
            } else {
                *self = below;                              /* :1296 */
                set_dirty(w, x, y);                         /* :1297 */
                if(rnd(&r, 1300) % 2 == 0) tile.moved = 1;
                *at(w, x, y + 1) = tile;                    /* :1306 */
                set_dirty(w, x, y + 1);                     /* :1307 */
                mark(c, x, y + 1);
            }
            if(rnd(&r, 1313) % 2 == 0) {                    /* :1313-1335 transmit movement */
                if(x > 0 && phys(w, at(w, x - 1, y + 1)) == FSS_PHYS_SAND) {
                    if(rnd(&r, 1316) % 2 == 0) at(w, x - 1, y + 1)->moved = 1;
                }
                if(x < w->width - 1 && phys(w, at(w, x + 1, y + 1)) == FSS_PHYS_SAND) {
                    if(rnd(&r, 1327) % 2 == 0) at(w, x + 1, y + 1)->moved = 1;
                }
            }
        }
    } else if(type == FSS_PHYS_SOUP) {                      /* :1338-1537 */
        float start, remaining, flow, dst;
        fss_cell bottom, left, right, top;
        int air_below, can_left, can_right;
        if(tile.fluid_amount == 0.0f) return;               /* :1344 */
        if(tile.fluid_amount < FLUID_MinValue) {            /* :1346-1350 */
            tile.fluid_amount = 0.0f;
            *self = tile;
            return;
        }
        if(particles && (double)tile.fluid_amount > 0.005 && phys_get(w, x, y + 1) == FSS_PHYS_AIR &&
           phys_get(w, x, y + 2) == FSS_PHYS_AIR && phys_get(w, x, y + 3) == FSS_PHYS_AIR &&
           phys_get(w, x, y + 4) == FSS_PHYS_AIR) {         /* :1352-1374 */
            int n = (int)(tile.fluid_amount / 4), i;
            *self = nothing();                              /* :1353 setTile: dirty */
            set_dirty(w, x, y);
            if(n < 1) n = 1;
            for(i = 0; i < n; i++) {
                fss_cell nt = liquid_seed(&tile);
                nt.fluid_amount = tile.fluid_amount / n;
                emit(c, x, y, 1366, r.cellkey, &nt);
            }
            return;
        }
        if(tile.moved) return;                              /* :1376 (moved == settled for liquids) */
        start = tile.fluid_amount;
        remaining = tile.fluid_amount;

        bottom = *at(w, x, y + 1);                          /* :1381 */
        air_below = phys(w, &bottom) == FSS_PHYS_AIR;
        if((air_below && c->iter <= 2) || bottom.material == tile.material) { /* :1384-1405 */
            dst = phys(w, &bottom) == FSS_PHYS_SOUP ? bottom.fluid_amount : 0.0f;
            flow = vertical_flow(start, dst) - dst;
            if(bottom.fluid_amount > 0 && flow > FLUID_MinFlow) flow *= FLUID_FlowSpeed;
            flow = clamp_flow(flow, start);
            if(flow != 0) {
                remaining -= flow;
                tile.fluid_amount_diff -= flow;
                if(phys(w, &bottom) == FSS_PHYS_AIR) *at(w, x, y + 1) = liquid_seed(&tile);
                at(w, x, y + 1)->fluid_amount_diff += flow;
            }
            w->flow_y[index] += flow;                       /* :1405 */
        } else if(c->iter == 0 && phys(w, &bottom) == FSS_PHYS_SOUP && bottom.material != tile.material) {
            if(rnd(&r, 1407) % 10 == 0) {                   /* :1406-1412 */
                *self = bottom;
                *at(w, x, y + 1) = tile;
                return;
            }
        }
        if(remaining < FLUID_MinValue) {                    /* :1414-1418 */
            tile.fluid_amount_diff -= remaining;
            *self = tile;
            return;
        }

        left = *at(w, x - 1, y);                            /* :1420-1424 */
        can_left = (phys(w, &left) == FSS_PHYS_AIR || left.material == tile.material) && !air_below;
        right = *at(w, x + 1, y);
        can_right = (phys(w, &right) == FSS_PHYS_AIR || right.material == tile.material) && !air_below;

        if(can_left) {                                      /* :1426-1448 */
            dst = phys(w, &left) == FSS_PHYS_SOUP ? left.fluid_amount : 0.0f;
            flow = (remaining - dst) / (can_right ? 3.0f : 2.0f);
            if(flow > FLUID_MinFlow) flow *= FLUID_FlowSpeed;
            flow = clamp_flow(flow, remaining);
            if(flow != 0) {
                remaining -= flow;
                tile.fluid_amount_diff -= flow;
                if(phys(w, &left) == FSS_PHYS_AIR) *at(w, x - 1, y) = liquid_seed(&tile);
                at(w, x - 1, y)->fluid_amount_diff += flow;
            }
            w->flow_x[index] -= flow;                       /* :1447 */
        }
        if(remaining < FLUID_MinValue) {
            tile.fluid_amount_diff -= remaining;
            *self = tile;
            return;
        }
        if(can_right) {                                     /* :1456-1478 (`canMoveLeft ? 2.0f : 2.0f`) */
            dst = phys(w, &right) == FSS_PHYS_SOUP ? right.fluid_amount : 0.0f;
            flow = (remaining - dst) / 2.0f;
            if(flow > FLUID_MinFlow) flow *= FLUID_FlowSpeed;
            flow = clamp_flow(flow, remaining);
            if(flow != 0) {
                remaining -= flow;
                tile.fluid_amount_diff -= flow;
                if(phys(w, &right) == FSS_PHYS_AIR) *at(w, x + 1, y) = liquid_seed(&tile);
                at(w, x + 1, y)->fluid_amount_diff += flow;
            }
            w->flow_x[index] += flow;                       /* :1477 */
        }
        if(remaining < FLUID_MinValue) {
            tile.fluid_amount_diff -= remaining;
            *self = tile;
            return;
        }
        top = *at(w, x, y - 1);                             /* :1486-1516 */
        if(phys(w, &top) == FSS_PHYS_AIR || top.material == tile.material) {
            dst = phys(w, &top) == FSS_PHYS_SOUP ? top.fluid_amount : 0.0f;
            flow = remaining - vertical_flow(remaining, dst);
            if(flow > FLUID_MinFlow) flow *= FLUID_FlowSpeed;
            flow = clamp_flow(flow, remaining);
            if(flow != 0) {
                remaining -= flow;
                tile.fluid_amount_diff -= flow;
                if(phys(w, &top) == FSS_PHYS_AIR) *at(w, x, y - 1) = liquid_seed(&tile);
                at(w, x, y - 1)->fluid_amount_diff += flow;
            }
            w->flow_y[index] -= flow;                       /* :1509 */
        } else if(c->iter == 0 && phys(w, &top) == FSS_PHYS_SOUP && top.material != tile.material) {
            if(rnd(&r, 1511) % 10 == 0) {
                *self = top;
                *at(w, x, y - 1) = tile;
                return;
            }
        }
        if(remaining < FLUID_MinValue) {
            tile.fluid_amount_diff -= remaining;
            *self = tile;
            return;
        }
        if(start == remaining) {                            /* :1524-1535 */
            tile.settle_count++;
            if(tile.settle_count >= 10) tile.moved = 1;
        } else {
            set_dirty(w, x, y);                             /* :1530 */
            if(phys(w, &top) == FSS_PHYS_SOUP) at(w, x, y - 1)->moved = 0;
            if(phys(w, &bottom) == FSS_PHYS_SOUP) at(w, x, y + 1)->moved = 0;
            if(phys(w, &left) == FSS_PHYS_SOUP) at(w, x - 1, y)->moved = 0;
            if(phys(w, &right) == FSS_PHYS_SOUP) at(w, x + 1, y)->moved = 0;
        }
        *self = tile;                                       /* :1537 */
    } else if(type == FSS_PHYS_GAS) {                       /* :1647-1663 */
        int above = phys(w, at(w, x, y - 1));
        int above_l = phys(w, at(w, x - 1, y - 1));
        int above_r = phys(w, at(w, x + 1, y - 1));
        if(above == FSS_PHYS_AIR && !((above_l == FSS_PHYS_AIR || above_r == FSS_PHYS_AIR) && rnd(&r, 1654) % 2 == 0)) {
            *self = *at(w, x, y - 1);
            *at(w, x, y - 1) = tile;
            set_dirty(w, x, y);                             /* :1656, :1659 */
            set_dirty(w, x, y - 1);
            mark(c, x, y - 1);
        }
    }
}

/* ---- scan 2, F/world.cpp:1669-1901 ----------------------------------------------------------- */
static void scan2_cell(chunk_t* c, int x, int y) {
    fsso_world* w = c->w;
    fss_cell* self = at(w, x, y);
    fss_cell tile;
    int type, dx = x - c->cx, dy = y - c->cy;
    rng_t r;
    if(c->vis[dy][dx]) return;                              /* :1676 */
    tile = *self;
    type = phys(w, &tile);
    r.cellkey = fss_rng_cell_key(c->tkey, (uint32_t)x, (uint32_t)y);
    r.n = 0;

    if(type == FSS_PHYS_SAND) {                             /* :1682-1807 */
        fss_cell bl = *at(w, x - 1, y + 1), br = *at(w, x + 1, y + 1);
        int can_l = can_move_into(w, &bl, &tile), can_r = can_move_into(w, &br, &tile);
        int stopped = !tile.moved;
        int slip = w->mat[tile.material].slipperyness;
        int should_move;
        if(stopped) {                                       /* :1697-1725 */
            int drop = 0, pil, unstable;
            for(pil = 0; pil < 10; pil++) {
                if(phys(w, at(w, x - 1, y + 1 + pil)) == FSS_PHYS_AIR || phys(w, at(w, x + 1, y + 1 + pil)) == FSS_PHYS_AIR) drop++;
            }
            unstable = drop + 1 - w->max_stability[tile.material];
            if(unstable > 0) {
                int chance = 1000 / unstable;
                if(chance < 1000) {
                    if(rnd(&r, 1715) % (uint32_t)chance == 0) {
                        stopped = 0;
                        self->moved = 1; /* the grid cell; the local copy `tile` keeps moved == false */
                    }
                }
            }
        }
        if(stopped || !(can_l || can_r)) {                  /* :1727-1734 */
            self->moved = 0;
            return;
        }
        should_move = rnd(&r, 1736) % (uint32_t)(2 * slip) != 0;
        if(should_move && (can_l || can_r)) {               /* :1738-1753 */
            if(rnd(&r, 1741) % 2 == 0) {
                if(phys(w, at(w, x, y + 1)) == FSS_PHYS_SAND) {
                    if(rnd(&r, 1744) % 2 == 0) at(w, x, y + 1)->moved = 1;
                }
            }
        }
        if(should_move && can_l && (!can_r || rnd(&r, 1755) % 2 == 0)) { /* :1755-1777 */
            if(phys(w, at(w, x - 1, y)) == FSS_PHYS_AIR) {
                *at(w, x - 1, y) = bl;
                set_dirty(w, x - 1, y);                     /* :1758 */
                mark(c, x - 1, y);
                *self = nothing();
            } else {
                *self = bl;
                c->vis[dy][dx] = 1;
            }
            set_dirty(w, x, y);                             /* :1761, :1764 */
            if(rnd(&r, 1768) % (uint32_t)(20 * slip) == 0) tile.moved = 0;
            *at(w, x - 1, y + 1) = tile;
            set_dirty(w, x - 1, y + 1);                     /* :1775 */
            mark(c, x - 1, y + 1);
        } else if(should_move && can_r) {                   /* :1778-1800: (x+1, y) is NOT marked visited */
            if(phys(w, at(w, x + 1, y)) == FSS_PHYS_AIR) {
                *at(w, x + 1, y) = br;
                set_dirty(w, x + 1, y);                     /* :1782 */
                *self = nothing();
            } else {
                *self = br;
                c->vis[dy][dx] = 1;
            }
            set_dirty(w, x, y);                             /* :1784, :1787 */
            if(rnd(&r, 1791) % (uint32_t)(20 * slip) == 0) tile.moved = 0;
            *at(w, x + 1, y + 1) = tile;
            set_dirty(w, x + 1, y + 1);                     /* :1798 */
            mark(c, x + 1, y + 1);
        } else {
            self->moved = 0;                                /* :1802 */
        }
    } else if(type == FSS_PHYS_SOUP) {                      /* :1808-1825 */
        tile.fluid_amount += tile.fluid_amount_diff;
        tile.fluid_amount_diff = 0.0f;
        if(tile.fluid_amount < FLUID_MinValue) *self = nothing();
        else *self = tile;
        set_dirty(w, x, y);                                 /* :1814, :1823: every liquid cell that scan 2 reaches */
        c->vis[dy][dx] = 1;
    } else if(type == FSS_PHYS_GAS) {                       /* :1879-1899 */
        int above_l = phys(w, at(w, x - 1, y - 1));
        int above_r = phys(w, at(w, x + 1, y - 1));
        if(above_l == FSS_PHYS_AIR && !(above_r == FSS_PHYS_AIR && rnd(&r, 1884) % 2 == 0)) {
            *self = *at(w, x - 1, y - 1);
            *at(w, x - 1, y - 1) = tile;
            set_dirty(w, x, y);                             /* :1886, :1889 */
            set_dirty(w, x - 1, y - 1);
            mark(c, x - 1, y - 1);
        } else if(above_r == FSS_PHYS_AIR) {
            *self = *at(w, x + 1, y - 1);
            *at(w, x + 1, y - 1) = tile;
            set_dirty(w, x, y);                             /* :1893, :1896 */
            set_dirty(w, x + 1, y - 1);
            mark(c, x + 1, y - 1);
        }
    }
}

/* ---- scan 3, F/world.cpp:1905-1974 ----------------------------------------------------------- */
static void scan3_cell(chunk_t* c, int x, int y) {
    fsso_world* w = c->w;
    fss_cell* self = at(w, x, y);
    fss_cell tile;
    int dx = x - c->cx, dy = y - c->cy;
    rng_t r;
    if(c->vis[dy][dx]) return;
    tile = *self;
    if(phys(w, &tile) != FSS_PHYS_GAS) return;              /* the SOUP branch is empty (:1918-1943) */
    r.cellkey = fss_rng_cell_key(c->tkey, (uint32_t)x, (uint32_t)y);
    r.n = 0;
    {
        int l = phys(w, at(w, x - 1, y)), rr = phys(w, at(w, x + 1, y));
        if(l == FSS_PHYS_AIR && !(rr == FSS_PHYS_AIR && rnd(&r, 1950) % 2 == 0)) {
            *self = *at(w, x - 1, y);
            *at(w, x - 1, y) = tile;
            set_dirty(w, x, y);                             /* :1952, :1955 */
            set_dirty(w, x - 1, y);
            mark(c, x - 1, y);
        } else if(rr == FSS_PHYS_AIR) {
            *self = *at(w, x + 1, y);
            *at(w, x + 1, y) = tile;
            set_dirty(w, x, y);                             /* :1959, :1962 */
            set_dirty(w, x + 1, y);
            mark(c, x + 1, y);
        } else if((w->mat[tile.material].flags & FSS_MAT_IS_STEAM) && w->condense_material >= 0) {
            if(rnd(&r, 1966) % 10 == 0) {
                *self = create_cell(w, w->condense_material, x, y, &r); /* Tiles::createWater() */
                set_dirty(w, x, y);                         /* :1968 */
            }
        }
    }
}

static void chunk_substep(fsso_world* w, uint32_t tick_ct, int iter, uint32_t tkey, int cx, int cy) {
    chunk_t c;
    int dx, dy;
    c.w = w;
    c.cx = cx;
    c.cy = cy;
    c.tick_ct = tick_ct;
    c.iter = iter;
    c.tkey = tkey;
    memset(c.vis, 0, sizeof c.vis);
    for(dy = CH - 1; dy >= 0; dy--)
        for(dx = 0; dx < CW; dx++) scan1_cell(&c, cx + dx, cy + dy);
    for(dy = CH - 1; dy >= 0; dy--)
        for(dx = 0; dx < CW; dx++) scan2_cell(&c, cx + dx, cy + dy);
    for(dy = CH - 1; dy >= 0; dy--)
        for(dx = 0; dx < CW; dx++) scan3_cell(&c, cx + dx, cy + dy);
}

/* F/world.cpp:1115-1141: phase tk activates chunks at zone + (tk%2, 1 - tk/2) * chunk + 2*chunk*(i, j) */
void fsso_substep_rows(fsso_world* w, uint32_t tick_ct, int iter, int tk, int cy_from, int cy_to) {
    int ofs_x = tk % 2, ofs_y = 1 - ((tk % 4) / 2);
    uint32_t tkey = fss_rng_tick_key(w->seed, tick_ct, (uint32_t)iter);
    int x_begin = w->zx + ofs_x * CW, x_end = w->zx + w->zw;
    int y_begin = w->zy + ofs_y * CH, y_end = w->zy + w->zh;
    int nx = x_begin < x_end ? (x_end - x_begin + 2 * CW - 1) / (2 * CW) : 0;
    int ny = y_begin < y_end ? (y_end - y_begin + 2 * CH - 1) / (2 * CH) : 0;
    int k;
#pragma omp parallel for schedule(dynamic, 64)
    for(k = 0; k < nx * ny; k++) {
        int cx = x_begin + (k % nx) * 2 * CW;
        int cy = y_begin + (k / nx) * 2 * CH;
        if(cy < cy_from || cy >= cy_to) continue;
        chunk_substep(w, tick_ct, iter, tkey, cx, cy);
    }
}

/* F/world.cpp:1108-1119, :2017 (the caller owns tickCt) */
void fsso_tick(fsso_world* w, uint32_t tick_ct) {
    int iter, tk;
    for(iter = 0; iter < 6; iter++)
        for(tk = 0; tk < 4; tk++) fsso_substep_rows(w, tick_ct, iter, tk, -0x40000000, 0x40000000);
}

/* ---- World::tickTemperature, F/world.cpp:2043-2143 ------------------------------------------- */
void fsso_tick_temperature_rows(fsso_world* w, int y_from, int y_to) {
    static const int NX[9] = {-1, -1, -1, 0, 0, 0, 1, 1, 1}; /* :2072-2116 order: x offset outer, y offset inner */
    static const int NY[9] = {-1, 0, 1, -1, 0, 1, -1, 0, 1};
    int x, y, i;
    int y0 = w->zy > y_from ? w->zy : y_from, y1 = (w->zy + w->zh) < y_to ? (w->zy + w->zh) : y_to;
#pragma omp parallel for private(x, i)
    for(y = y0; y < y1; y++) {
        for(x = w->zx; x < w->zx + w->zw; x++) {
            float n = 0.01f, v = 0.0f;
            const fss_cell* self = at(w, x, y);
            const fss_material* m = &w->mat[self->material];
            int32_t out;
            for(i = 0; i < 9; i++) {
                const fss_cell* c = at(w, x + NX[i], y + NY[i]);
                if(c->temperature) {
                    float factor = abs(c->temperature) / 64.0f * w->mat[c->material].conduction_other;
                    v += c->temperature * factor;
                    n += factor;
                }
            }
            if(v != 0) {                                    /* :2128-2132; addTemp is Uint32 */
                out = (int32_t)(m->add_temp + (v / n * m->conduction_self) + (self->temperature * (1 - m->conduction_self)));
            } else {
                out = (int32_t)(m->add_temp + (uint32_t)self->temperature);
            }
            w->new_temps[(size_t)y * w->width + x] = out;
        }
    }
#pragma omp parallel for private(x)
    for(y = y0; y < y1; y++)                                /* :2137-2141 */
        for(x = w->zx; x < w->zx + w->zw; x++) at(w, x, y)->temperature = w->new_temps[(size_t)y * w->width + x];
}

void fsso_tick_temperature(fsso_world* w) { fsso_tick_temperature_rows(w, -0x40000000, 0x40000000); }

/* ---- bookkeeping ----------------------------------------------------------------------------- */
fsso_world* fsso_create(int width, int height, uint32_t seed, uint32_t flags, int condense_material) {
    fsso_world* w = (fsso_world*)calloc(1, sizeof *w);
    size_t n = (size_t)width * height, i;
    w->width = width;
    w->height = height;
    w->seed = seed;
    w->flags = flags;
    w->condense_material = condense_material;
    w->cells = (fss_cell*)malloc(n * sizeof(fss_cell));
    w->new_temps = (int32_t*)calloc(n, sizeof(int32_t));
    w->flow_x = (float*)calloc(n, sizeof(float));
    w->flow_y = (float*)calloc(n, sizeof(float));
    w->dirty = (uint8_t*)calloc(n, 1);
    w->prev_flow_x = (float*)calloc(n, sizeof(float));
    w->prev_flow_y = (float*)calloc(n, sizeof(float));
    memset(w->alpha, 255, sizeof w->alpha);
    for(i = 0; i < n; i++) w->cells[i] = nothing();
    w->fire_id = -1;
    return w;
}

void fsso_destroy(fsso_world* w) {
    int i;
    if(!w) return;
    for(i = 0; i < FSS_MAX_TEXTURES; i++) free(w->tex[i].px);
    free(w->cells);
    free(w->new_temps);
    free(w->flow_x);
    free(w->flow_y);
    free(w->dirty);
    free(w->prev_flow_x);
    free(w->prev_flow_y);
    free(w->parts);
    free(w->plist);
    free(w);
}

int fsso_set_materials(fsso_world* w, const fss_material* table, int n) {
    int i;
    if(n < 1 || n > FSS_MAX_MATERIALS) return FSS_ERR_INVALID;
    memcpy(w->mat, table, (size_t)n * sizeof(fss_material));
    w->n_mat = n;
    w->fire_id = -1;
    for(i = 0; i < n; i++) {
        int slip = table[i].slipperyness;
        /* `int maxStability = 8 / sqrt(slipperyness) + 1;`  F/world.cpp:1710 (double arithmetic) */
        w->max_stability[i] = slip > 0 ? (int)(8 / sqrt((double)slip) + 1) : 0x3fffffff;
        if((table[i].flags & FSS_MAT_IS_FIRE) && w->fire_id < 0) w->fire_id = i;
        if(table[i].physics_type == FSS_PHYS_SAND && slip < 1) return FSS_ERR_INVALID;
    }
    return FSS_OK;
}

int fsso_set_texture(fsso_world* w, int index, const uint32_t* argb, int tw, int th) {
    if(index < 0 || index >= FSS_MAX_TEXTURES || tw < 1 || th < 1) return FSS_ERR_INVALID;
    free(w->tex[index].px);
    w->tex[index].px = (uint32_t*)malloc((size_t)tw * th * 4);
    memcpy(w->tex[index].px, argb, (size_t)tw * th * 4);
    w->tex[index].w = tw;
    w->tex[index].h = th;
    return FSS_OK;
}

int fsso_set_tick_zone(fsso_world* w, int x, int y, int zw, int zh) {
    if(x % FSS_ZONE_ALIGN_X || zw % FSS_ZONE_ALIGN_X || zh % FSS_ZONE_ALIGN_Y) return FSS_ERR_ALIGN;
    if(x < FSS_ZONE_MARGIN || y < FSS_ZONE_MARGIN || x + zw > w->width - FSS_ZONE_MARGIN || y + zh > w->height - FSS_ZONE_MARGIN)
        return FSS_ERR_ALIGN;
    w->zx = x;
    w->zy = y;
    w->zw = zw;
    w->zh = zh;
    return FSS_OK;
}

fss_cell* fsso_cells(fsso_world* w) { return w->cells; }
float* fsso_flow_x(fsso_world* w) { return w->flow_x; }
float* fsso_flow_y(fsso_world* w) { return w->flow_y; }
uint8_t* fsso_dirty(fsso_world* w) { return w->dirty; }

int fsso_set_material_looks(fsso_world* w, const uint8_t* alpha, const uint32_t* emit_color, int n) {
    if(n < 0 || n > FSS_MAX_MATERIALS) return FSS_ERR_INVALID;
    memcpy(w->alpha, alpha, (size_t)n);
    memcpy(w->emit_color, emit_color, (size_t)n * 4);
    return FSS_OK;
}

/* The `dirty` job of Game::tick, F/Game.cpp:2579-2650, and the memset that follows it (:2751).  pixels[k] = row-major
 * width * height * 4 bytes: 0 main layer, 1 fire, 2 emission, 3 flow.  moving[m] counts the dirty cells of material m
 * (the reference's counters are uint16_t and wrap, F/Game.hpp:204; these do not). */
void fsso_render_dirty(fsso_world* w, uint8_t* const pixels[4], uint64_t* moving, int n_materials, fss_render_info* info) {
    size_t i, n = (size_t)w->width * w->height;
    uint8_t *px = pixels[0], *fire = pixels[1], *emi = pixels[2], *flw = pixels[3];
    int m;
    memset(info, 0, sizeof *info);
    for(m = 0; m < n_materials; m++) moving[m] = 0;
    for(i = 0; i < n; i++) {
        const size_t o = i * 4;
        const fss_cell* t = &w->cells[i];
        const fss_material* mat = &w->mat[t->material];
        if(!w->dirty[i]) continue;
        info->had_dirty = 1;
        if(t->material < n_materials) moving[t->material]++;
        if(mat->physics_type == FSS_PHYS_AIR) {                     /* :2590-2607 */
            memset(px + o, 0, 4);
            memset(fire + o, 0, 4);
            memset(emi + o, 0, 4);
            w->flow_y[i] = 0;
            w->flow_x[i] = 0;
        } else {
            uint32_t color = t->color, emit = w->emit_color[t->material];
            px[o + 2] = color & 0xff;                                /* :2611-2614 */
            px[o + 1] = (color >> 8) & 0xff;
            px[o + 0] = (color >> 16) & 0xff;
            px[o + 3] = w->alpha[t->material];
            emi[o + 2] = emit & 0xff;                                /* :2616-2619 */
            emi[o + 1] = (emit >> 8) & 0xff;
            emi[o + 0] = (emit >> 16) & 0xff;
            emi[o + 3] = (emit >> 24) & 0xff;
            if(mat->flags & FSS_MAT_IS_FIRE) {                       /* :2621-2627 */
                fire[o + 2] = color & 0xff;
                fire[o + 1] = (color >> 8) & 0xff;
                fire[o + 0] = (color >> 16) & 0xff;
                fire[o + 3] = w->alpha[t->material];
                info->had_fire = 1;
            }
            if(mat->physics_type == FSS_PHYS_SOUP) {                 /* :2628-2643: double arithmetic, like the reference */
                float newFlowX = w->prev_flow_x[i] + (w->flow_x[i] - w->prev_flow_x[i]) * 0.25;
                float newFlowY = w->prev_flow_y[i] + (w->flow_y[i] - w->prev_flow_y[i]) * 0.25;
                double gx, gy;
                if(newFlowY < 0) newFlowY *= 0.5;
                gy = newFlowY * (3.0 / mat->iterations + 0.5) / 4.0 + 0.5;
                gx = newFlowX * (3.0 / mat->iterations + 0.5) / 4.0 + 0.5;
                gy = gy < 0.0 ? 0.0 : gy; /* std::max(v, 0.0) = (v < 0.0) ? 0.0 : v */
                gx = gx < 0.0 ? 0.0 : gx;
                gy = 1.0 < gy ? 1.0 : gy; /* std::min(v, 1.0) = (1.0 < v) ? 1.0 : v */
                gx = 1.0 < gx ? 1.0 : gx;
                flw[o + 2] = 0;
                flw[o + 1] = (uint8_t)(gy * 255);
                flw[o + 0] = (uint8_t)(gx * 255);
                flw[o + 3] = 0xff;
                info->had_flow = 1;
                w->prev_flow_x[i] = newFlowX;
                w->prev_flow_y[i] = newFlowY;
            }
            w->flow_y[i] = 0;                                        /* :2642-2647 */
            w->flow_x[i] = 0;
        }
    }
    if(info->had_dirty) memset(w->dirty, 0, n);                      /* :2751 */
}
size_t fsso_num_particles(fsso_world* w) { return w->n_parts; }

/* the order World::tick appends to World::particles, see fss_spawn_order_key (F/world.cpp:1140-1157, :1989-1994) */
static int g_cmp_zx, g_cmp_zy;
static int cmp_part(const void* a, const void* b) {
    const fss_particle_spawn* p = (const fss_particle_spawn*)a;
    const fss_particle_spawn* q = (const fss_particle_spawn*)b;
    uint64_t kp = fss_spawn_order_key(p->tick_iter, p->x, p->y, g_cmp_zx, g_cmp_zy);
    uint64_t kq = fss_spawn_order_key(q->tick_iter, q->x, q->y, g_cmp_zx, g_cmp_zy);
    return kp < kq ? -1 : kp > kq ? 1 : 0;
}

static void sort_spawned(fsso_world* w) {
    g_cmp_zx = w->zx;
    g_cmp_zy = w->zy;
    qsort(w->parts, w->n_parts, sizeof(fss_particle_spawn), cmp_part); /* equal keys are identical records */
}

size_t fsso_drain_particles(fsso_world* w, fss_particle_spawn* out, size_t cap) {
    size_t n = w->n_parts < cap ? w->n_parts : cap;
    sort_spawned(w);
    memcpy(out, w->parts, n * sizeof(fss_particle_spawn));
    w->n_parts = 0;
    return n;
}

/* ---- World::particles, F/world.cpp:2171-2311 ----------------------------------------------------------- */
static void plist_reserve(fsso_world* w, size_t extra) {
    if(w->n_plist + extra > w->cap_plist) {
        while(w->n_plist + extra > w->cap_plist) w->cap_plist = w->cap_plist ? w->cap_plist * 2 : 1024;
        w->plist = (fss_particle*)realloc(w->plist, w->cap_plist * sizeof(fss_particle));
    }
}

static void adopt_spawned(fsso_world* w);

/* a host-side particles.push_back: lands behind everything the ticks so far appended */
void fsso_particles_add(fsso_world* w, const fss_particle* p, size_t n) {
    adopt_spawned(w);
    plist_reserve(w, n);
    memcpy(w->plist + w->n_plist, p, n * sizeof(fss_particle));
    w->n_plist += n;
}

fss_particle* fsso_particles(fsso_world* w, size_t* n) {
    *n = w->n_plist;
    return w->plist;
}

/* `new Particle(...)` at F/world.cpp:1180-1183, :1291, :1366 (the DO_MULTITHREADING branches, F/world.cpp:1097).  The two
 * rand() calls of one constructor expression: g++ evaluates the arguments right to left, so vy draws first. */
static void particle_from_spawn(const fss_particle_spawn* s, uint32_t j, fss_particle* p) {
    uint32_t ry = fss_rand(s->cell_key, s->site, 2u * j), rx = fss_rand(s->cell_key, s->site, 2u * j + 1u);
    memset(p, 0, sizeof *p);
    p->tile = s->cell;
    p->x = (float)(int)s->x;
    p->fade_time = 60; /* F/Particle.hpp:27 */
    if(s->site == 1180) {
        p->y = (float)((int)s->y - 1);
        p->vx = (float)((int)(rx % 10) - 5) / 20.0f;
        p->vy = -((float)(int)(ry % 10) / 10.0f) / 3.0f + -0.5f;
        p->ay = 0.01f;
        p->temporary = 1;
        p->lifetime = 30;
        p->fade_time = 10;
    } else if(s->site == 1291) {
        p->y = (float)((int)s->y + 1);
        p->vx = (float)((int)(rx % 10) - 5) / 20.0f;
        p->vy = (float)(-((int)(ry % 2) + 3)) / 10.0f + 1.5f;
        p->ay = 0.1f;
    } else {
        p->y = (float)((int)s->y + 1);
        p->vx = (float)((int)(rx % 10) - 5) / 30.0f;
        p->vy = (float)(-((int)(ry % 2) + 3)) / 10.0f + 1.0f;
        p->ay = 0.1f;
    }
}

/* what World::tick's `particles.insert(...)` did for the ticks since the last call */
static void adopt_spawned(fsso_world* w) {
    size_t i;
    uint32_t j = 0;
    sort_spawned(w);
    plist_reserve(w, w->n_parts);
    for(i = 0; i < w->n_parts; i++) {
        const fss_particle_spawn* s = &w->parts[i];
        j = (i > 0 && s[-1].tick_iter == s->tick_iter && s[-1].x == s->x && s[-1].y == s->y) ? j + 1 : 0;
        particle_from_spawn(s, j, &w->plist[w->n_plist++]);
    }
    w->n_parts = 0;
}

/* the lambda of the remove_if, F/world.cpp:2179-2310; returns 1 when the particle leaves the list */
static int particle_step(fsso_world* w, fss_particle* cur) {
    int lx, ly, div, i;
    float dvx, dvy;
    if(cur->temporary && cur->lifetime <= 0) return 1; /* :2182-2186 */
    if(cur->target_force != 0) {                       /* :2188-2201 */
        float tdx = cur->target_x - cur->x;
        float tdy = cur->target_y - cur->y;
        float normFac = sqrtf(tdx * tdx + tdy * tdy);
        cur->vx += tdx / normFac * cur->target_force;
        cur->vy += tdy / normFac * cur->target_force;
        if(normFac < 100) {
            cur->vx *= 0.95f;
            cur->vy *= 0.95f;
        }
    }
    lx = (int)cur->x;
    ly = (int)cur->y;
    if(cur->x < 0 || (int)(cur->x) >= w->width || cur->y < 0 || (int)(cur->y) >= w->height) return 1; /* :2206-2209 */
    if(!(lx >= w->zx && ly >= w->zy && lx < w->zx + w->zw && ly < w->zy + w->zh)) return 0;               /* :2211 */
    cur->vx += cur->ax;
    cur->vy += cur->ay;
    div = (int)((fabsf(cur->vx) + fabsf(cur->vy)) + 1); /* :2216 */
    dvx = cur->vx / div;
    dvy = cur->vy / div;
    for(i = 0; i < div; i++) {
        fss_cell* hit;
        cur->x += dvx;
        cur->y += dvy;
        if(cur->x < 0 || (int)(cur->x) >= w->width || cur->y < 0 || (int)(cur->y) >= w->height) return 1; /* :2225-2228 */
        hit = at(w, (int)cur->x, (int)cur->y);
        if(!cur->phase && phys(w, hit) != FSS_PHYS_AIR) { /* :2230 */
            int isObject = phys(w, hit) == FSS_PHYS_PASSABLE; /* PhysicsType::OBJECT */
            switch(cur->in_object_state) {                    /* :2234-2246 */
            case 0: cur->in_object_state = isObject ? 1 : 2; break;
            case 1: if(!isObject) cur->in_object_state = 2; break;
            }
            if(!isObject || cur->in_object_state == 2) {
                if(cur->temporary) return 1; /* :2249-2253 */
                if(phys(w, at(w, lx, ly)) != FSS_PHYS_AIR) {
                    /* square spiral around the particle, first AIR cell (or, for a liquid, first cell of its own
                     * material) takes it: :2263-2290 */
                    int succeeded = 0;
                    int X = 32, Y = 32, x = 0, y = 0, dx = 0, dy = -1, t = X > Y ? X : Y, maxI = t * t, j;
                    for(j = 0; j < maxI; j++) {
                        if((-X / 2 <= x) && (x <= X / 2) && (-Y / 2 <= y) && (y <= Y / 2)) {
                            int px = (int)(cur->x + x), py = (int)(cur->y + y);
                            /* the reference indexes tiles[] unchecked here; outside the world we skip (fss.h) */
                            if(px >= 0 && px < w->width && py >= 0 && py < w->height) {
                                fss_cell* c = at(w, px, py);
                                if(phys(w, c) == FSS_PHYS_AIR) {
                                    *c = cur->tile;
                                    set_dirty(w, px, py);   /* :2269 */
                                    succeeded = 1;
                                    break;
                                } else if(w->mat[cur->tile.material].physics_type == FSS_PHYS_SOUP && cur->tile.material == c->material) {
                                    c->fluid_amount += cur->tile.fluid_amount;
                                    set_dirty(w, px, py);   /* :2275 */
                                    succeeded = 1;
                                    break;
                                }
                            }
                        }
                        if((x == y) || ((x < 0) && (x == -y)) || ((x > 0) && (x == 1 - y))) {
                            t = dx;
                            dx = -dy;
                            dy = t;
                        }
                        x += dx;
                        y += dy;
                    }
                    if(succeeded) return 1;
                    cur->vy = -4; /* :2296-2299 */
                    cur->y -= 16;
                    return 0;
                } else {
                    *at(w, lx, ly) = cur->tile; /* :2301-2306 */
                    set_dirty(w, lx, ly);
                    return 1;
                }
            }
        }
    }
    if(cur->lifetime > 0) cur->lifetime--; /* :2311-2313 */
    return 0;
}

void fsso_tick_particles(fsso_world* w) {
    size_t i, out = 0;
    adopt_spawned(w);
    for(i = 0; i < w->n_plist; i++) {
        if(!particle_step(w, &w->plist[i])) {
            if(out != i) w->plist[out] = w->plist[i];
            out++;
        }
    }
    w->n_plist = out;
}

uint64_t fsso_hash_rows(const fss_cell* cells, int width, int x0, int y0, int w, int h, int global_y0) {
    uint64_t sum = 0;
    int x, y;
    for(y = y0; y < y0 + h; y++) {
        for(x = x0; x < x0 + w; x++) {
            const fss_cell* c = &cells[(size_t)y * width + x];
            uint32_t a, d, mms = (uint32_t)c->material | ((uint32_t)(c->moved != 0) << 16) | ((uint32_t)c->settle_count << 24);
            memcpy(&a, &c->fluid_amount, 4);
            memcpy(&d, &c->fluid_amount_diff, 4);
            sum += fss_cell_digest((uint32_t)x, (uint32_t)(y + global_y0), mms, c->color, (uint32_t)c->temperature, a, d);
        }
    }
    return sum;
}

void fsso_histogram(const fss_cell* cells, size_t n, uint64_t* counts, int n_materials) {
    size_t i;
    memset(counts, 0, (size_t)n_materials * sizeof(uint64_t));
    for(i = 0; i < n; i++)
        if(cells[i].material < n_materials) counts[cells[i].material]++;
}
