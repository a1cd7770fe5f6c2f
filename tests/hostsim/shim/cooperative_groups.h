/* tests/hostsim: stand-in for <cooperative_groups.h>; the cooperative-launch kernel itself is not run on the host (its
 * barriers need real threads), only the per-particle functions it shares with the multi-kernel path. */
#pragma once
namespace cooperative_groups {
struct grid_group {
    void sync() const {}
};
static inline grid_group this_grid() { return grid_group(); }
} // namespace cooperative_groups
