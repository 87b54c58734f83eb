// fluid_solver.hpp -- the solver interface of the reference (src/fluid_solver.hpp:8-25), re-declared
// API-compatibly: one pure virtual solve() taking six host grids and three scalars; density and the
// two velocity grids are updated in place, the source grids are const.  In a reference checkout the
// reference's own header is used instead; this copy makes the repository
// self-contained for the headless driver and the tests.
#pragma once

#include "grid.hpp"

class fluid_solver {
public:
    virtual ~fluid_solver() = default;

    virtual void solve(grid<float>& density_grid,
                       grid<float> const& density_source_grid,
                       float const diffusion_rate,
                       grid<float>& horizontal_velocity_grid,
                       grid<float>& vertical_velocity_grid,
                       grid<float> const& horizontal_velocity_source_grid,
                       grid<float> const& vertical_velocity_source_grid,
                       float const viscosity,
                       float const dt) = 0;
};
