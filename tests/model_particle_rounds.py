"""A CPU model of the round scheme of csrc/fss_particles.cuh (speculate / cascade to a fixed point / commit), to check the
ARGUMENT for its exactness apart from the CUDA code: whatever the order in which the threads of a pass run and whenever
their table updates become visible to each other, and with hash tables so small that they collide all the time, the rounds
must (a) retire at least one particle each and (b) end with the grid and the list of the serial loop.

The GPU parity tests only see the schedules the hardware happened to produce; this model enumerates hostile ones.  It
restates the scheme, not the kernels: a stand-alone, simplified tickParticles (no target force, no objects, integer
velocities are enough) with the same read / write structure: a path of cells, a hit, the start-cell test, a spiral landing
search with `first free cell, or first cell of my own liquid`, a bounce.  Run by tests/test_particle_round_model.py.
"""
import numpy as np

AIR, SOLID, LIQ = 0, 1, 2  # cell kinds; a LIQ cell carries an amount
TILE = 8
SPIRAL_R = 3  # the model's landing search covers (2 * SPIRAL_R + 1)^2 cells


def spiral(r=SPIRAL_R):
    """square spiral like world.cpp:2256-2290: (0,0), then rings outwards"""
    out = []
    x = y = 0
    dx, dy = 0, -1
    for _ in range((2 * r + 2) ** 2):
        if -r <= x <= r and -r <= y <= r:
            out.append((x, y))
        if x == y or (x < 0 and x == -y) or (x > 0 and x == 1 - y):
            dx, dy = -dy, dx
        x, y = x + dx, y + dy
    return out


SPIRAL = spiral()


class Particle:
    __slots__ = ("x", "y", "vx", "vy", "kind", "amount", "temporary")

    def __init__(self, x, y, vx, vy, kind, amount, temporary):
        self.x, self.y, self.vx, self.vy, self.kind, self.amount, self.temporary = x, y, vx, vy, kind, amount, temporary

    def copy(self):
        return Particle(self.x, self.y, self.vx, self.vy, self.kind, self.amount, self.temporary)

    def key(self):
        return (self.x, self.y, self.vx, self.vy, self.kind, self.amount, self.temporary)


def path_positions(p):
    """the sub-steps of one tick: |vx| + |vy| + 1 equal integer-ish steps (the model moves on a lattice along a staircase)"""
    div = abs(p.vx) + abs(p.vy) + 1
    pts = []
    for i in range(1, div + 1):
        pts.append((p.x + (p.vx * i) // div, p.y + (p.vy * i) // div))
    return pts


def walk(kind_grid, p, look=None):
    """one particle's tick against a grid, WITHOUT writing: returns (action, target, new particle state, hit step).
    action: 'keep' | 'remove' | 'deposit' | 'merge'.  `look(x, y)` is called for every cell whose kind is read."""
    h, w = kind_grid.shape
    q = p.copy()
    sx, sy = p.x, p.y
    pts = path_positions(p)
    for step, (x, y) in enumerate(pts, 1):
        if x < 0 or x >= w or y < 0 or y >= h:
            return "remove", None, q, 0
        if look:
            look(x, y)
        if kind_grid[y, x] != AIR:
            if p.temporary:
                return "remove", None, q, step
            if look:
                look(sx, sy)
            if kind_grid[sy, sx] != AIR:
                for ox, oy in SPIRAL:
                    px, py = x + ox, y + oy
                    if px < 0 or px >= w or py < 0 or py >= h:
                        continue
                    if look:
                        look(px, py)
                    if kind_grid[py, px] == AIR:
                        return "deposit", (px, py), q, step
                    if p.kind == LIQ and kind_grid[py, px] == LIQ:
                        return "merge", (px, py), q, step
                q.x, q.y, q.vy = x, y - 4, -2  # bounce
                return "keep", None, q, step
            return "deposit", (sx, sy), q, step
    q.x, q.y = pts[-1]
    q.vy += 1
    return "keep", None, q, 0


def apply(kind_grid, amount_grid, p, action, target):
    if action == "deposit":
        kind_grid[target[1], target[0]] = p.kind
        amount_grid[target[1], target[0]] = p.amount
    elif action == "merge":
        amount_grid[target[1], target[0]] = np.float32(amount_grid[target[1], target[0]] + p.amount)  # order-sensitive in fp32


def serial(kind_grid, amount_grid, parts):
    """the reference's remove_if loop"""
    kg, ag = kind_grid.copy(), amount_grid.copy()
    out = []
    for p in parts:
        action, target, q, _ = walk(kg, p)
        apply(kg, ag, p, action, target)
        if action == "keep":
            out.append(q)
    return kg, ag, out


class Tables:
    """min-rank tables, hashed on purpose into very few slots"""

    def __init__(self, slots, tile_slots):
        self.slots, self.tile_slots = slots, tile_slots
        self.clear()

    def clear(self):
        inf = 1 << 30
        self.w_act = [inf] * self.slots
        self.r_act = [inf] * self.slots
        self.w_pot = [inf] * self.slots
        self.r_pot = [inf] * self.slots
        self.box = [inf] * self.tile_slots

    def h(self, x, y):
        return (x * 7919 + y * 104729) % self.slots

    def t(self, x, y):
        return ((x // TILE) * 31 + (y // TILE) * 17) % self.tile_slots


def reach_tiles(kind_grid, p, T):
    """every tile the particle's path and its landing searches can touch"""
    h, w = kind_grid.shape
    pts = [(p.x, p.y)] + path_positions(p)
    x0 = max(0, min(x for x, _ in pts) - SPIRAL_R)
    x1 = min(w - 1, max(x for x, _ in pts) + SPIRAL_R)
    y0 = max(0, min(y for _, y in pts) - SPIRAL_R)
    y1 = min(h - 1, max(y for _, y in pts) + SPIRAL_R)
    return {T.t(x, y) for x in range(x0, x1 + 1) for y in range(y0, y1 + 1)}


def potential_sets(kind_grid, p, k0, rank, T, pending):
    """where an unsure particle can still look (R) and land (W): csrc/fss_particles.cuh PotentialSets.  `pending` collects the
    table updates (table, slot) so that the caller decides when they become visible."""
    h, w = kind_grid.shape
    pts = path_positions(p)

    def lower_writer(x, y):
        return T.w_act[T.h(x, y)] < rank or T.w_pot[T.h(x, y)] < rank or T.box[T.t(x, y)] < rank

    sx, sy = p.x, p.y
    pending.append(("r_pot", T.h(sx, sy)))
    kmax = k0 if k0 else len(pts)
    if p.temporary:
        for (x, y) in pts[:kmax]:
            if x < 0 or x >= w or y < 0 or y >= h:
                break
            pending.append(("r_pot", T.h(x, y)))
        return
    s_free = kind_grid[sy, sx] == AIR
    if s_free:
        pending.append(("w_pot", T.h(sx, sy)))
    search = (not s_free) or lower_writer(sx, sy)
    for step, (x, y) in enumerate(pts[:kmax], 1):
        if x < 0 or x >= w or y < 0 or y >= h:
            break
        pending.append(("r_pot", T.h(x, y)))
        if not search or not (step == k0 or lower_writer(x, y)):
            continue
        for ox, oy in SPIRAL:
            px, py = x + ox, y + oy
            if px < 0 or px >= w or py < 0 or py >= h:
                continue
            pending.append(("r_pot", T.h(px, py)))
            if kind_grid[py, px] == AIR:
                pending.append(("w_pot", T.h(px, py)))
                if not lower_writer(px, py):
                    break
            elif p.kind == LIQ and kind_grid[py, px] == LIQ:
                pending.append(("w_pot", T.h(px, py)))
                break


def rounds(kind_grid, amount_grid, parts, rng, slots=61, tile_slots=7, visibility="random", budget=1 << 30, wake=True):
    """the round scheme.  visibility: 'late' = table updates of a pass become visible only after the pass (all threads read
    the old tables), 'now' = immediately (threads run one after the other in a random order), 'random' = each update flips a
    coin.  budget: potential sets longer than this fall back to the reach box.  wake: a particle only looks at the tables again
    when somebody whose reach shares a tile with its own has published something since it last looked (the kernels'
    `w_when` / `ran`).  Returns (kind grid, amount grid, surviving
    list, number of rounds)."""
    kg, ag = kind_grid.copy(), amount_grid.copy()
    n = len(parts)
    decided = [False] * n
    keep = [False] * n
    out = [None] * n
    T = Tables(slots, tile_slots)
    n_rounds = 0
    while not all(decided):
        n_rounds += 1
        T.clear()
        act = [None] * n
        seen = [None] * n
        sure = [True] * n
        # speculate
        for i in rng.permutation(n):
            if decided[i]:
                continue
            cells = []
            action, target, q, k0 = walk(kg, parts[i], lambda x, y: cells.append((x, y)))
            act[i], out[i], seen[i] = (action, target, k0), q, cells
            if not cells:
                decided[i], keep[i] = True, action == "keep"
                continue
            for (x, y) in cells:
                T.r_act[T.h(x, y)] = min(T.r_act[T.h(x, y)], i)
            if target is not None:
                T.w_act[T.h(*target)] = min(T.w_act[T.h(*target)], i)
        # cascade to a fixed point
        listed = [0] * n
        boxed = [False] * n
        when = [0] * tile_slots  # last pass in which a particle that can reach the tile published something (0 = speculation)
        ran = [0] * n
        reach = [None if decided[i] else reach_tiles(kg, parts[i], T) for i in range(n)]
        for pno in range(1, 10 * n + 11):
            changed = False
            late = []
            late_when = []
            for i in rng.permutation(n):
                if decided[i]:
                    continue
                if boxed[i]:
                    continue
                if wake:
                    if max(when[t] for t in reach[i]) < ran[i]:
                        continue
                    ran[i] = pno
                if sure[i]:
                    hit = any(T.w_act[T.h(x, y)] < i or T.w_pot[T.h(x, y)] < i or T.box[T.t(x, y)] < i for (x, y) in seen[i])
                    if not hit:
                        continue
                    sure[i] = False
                    changed = True
                pend = []
                potential_sets(kg, parts[i], act[i][2], i, T, pend)
                if len(pend) > budget:  # too much to list: the whole reach, by tiles
                    pend = [("box", t) for t in reach_tiles(kg, parts[i], T)]
                    boxed[i] = True
                if len(pend) > listed[i]:
                    listed[i] = len(pend)
                    changed = True
                    for t in reach[i]:
                        if visibility == "now" or (visibility == "random" and rng.random() < 0.5):
                            when[t] = max(when[t], pno)
                        else:
                            late_when.append(t)
                for (name, slot) in pend:
                    tab = getattr(T, name)
                    if tab[slot] > i:
                        if visibility == "now" or (visibility == "random" and rng.random() < 0.5):
                            tab[slot] = i
                        else:
                            late.append((name, slot, i))
                        changed = True
            for (name, slot, i) in late:
                tab = getattr(T, name)
                tab[slot] = min(tab[slot], i)
            for t in late_when:
                when[t] = max(when[t], pno)
            if not changed:
                break
        else:
            raise AssertionError("cascade did not reach a fixed point")
        # commit
        committed = 0
        go = []
        for i in range(n):
            if decided[i] or not sure[i]:
                continue
            action, target, _ = act[i]
            if target is not None:
                hh = T.h(*target)
                if T.r_act[hh] < i or T.r_pot[hh] < i or T.box[T.t(*target)] < i:
                    continue
            go.append(i)
        for i in rng.permutation(len(go)):  # in any order: that is the claim
            j = go[i]
            action, target, _ = act[j]
            apply(kg, ag, parts[j], action, target)
            decided[j], keep[j] = True, action == "keep"
            committed += 1
        assert committed > 0 or all(decided), f"round {n_rounds} retired nobody"
    return kg, ag, [out[i] for i in range(n) if keep[i]], n_rounds


def random_scene(rng, size=40, n=120):
    kind = np.zeros((size, size), dtype=np.int8)
    amount = np.zeros((size, size), dtype=np.float32)
    r = rng.random((size, size))
    kind[r < 0.25] = SOLID
    kind[(r >= 0.25) & (r < 0.40)] = LIQ
    amount[kind == LIQ] = rng.random((size, size)).astype(np.float32)[kind == LIQ]
    kind[size - 6:, :] = SOLID  # a floor to land on
    parts = []
    for _ in range(n):
        parts.append(Particle(int(rng.integers(0, size)), int(rng.integers(0, size - 8)), int(rng.integers(-2, 3)), int(rng.integers(0, 5)),
                              LIQ if rng.random() < 0.4 else SOLID, np.float32(rng.random()), bool(rng.random() < 0.1)))
    return kind, amount, parts
