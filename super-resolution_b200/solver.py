"""Host mirror of the reference's solver front end for the device-resident solver.

`IrlsMapSolverOptions` carries the fields and defaults of MapSolverOptions / IRLSMapSolverOptions
(map_solver.h:25-62, irls_map_solver.h:20-36); `solve()` is IRLSMapSolver::Solve
(irls_map_solver.cpp:192-265): thresholds scaled by num_parameters * sum(lambda) when that exceeds 1
(:161-171, map_solver.cpp:16-26), one solver round per channel when split_channels is set, each
round one srb_solve_irls call -- conjugate gradients (ALGLIB mincg restated, csrc/srb_cg.h) and the
IRLS re-weighting with every vector on the device."""
from dataclasses import dataclass, replace

import numpy as np


@dataclass
class IrlsMapSolverOptions:
    max_num_solver_iterations: int = 50            # map_solver.h:54
    gradient_norm_threshold: float = 1.0e-6        # :58
    cost_decrease_threshold: float = 1.0e-6        # :60
    parameter_variation_threshold: float = 1.0e-6  # :62
    split_channels: bool = False
    least_squares_solver: str = "cg"               # CG_SOLVER (default) | "lbfgs" (map_solver.h:20-23)
    num_lbfgs_hessian_corrections: int = 5         # :51
    max_num_irls_iterations: int = 20              # irls_map_solver.h:27
    irls_cost_difference_threshold: float = 1.0e-5  # :35

    def adjusted(self, num_parameters, regularization_parameter_sum):
        """AdjustThresholdsAdaptively: scale the thresholds up (never down)."""
        scale = num_parameters * regularization_parameter_sum
        if scale < 1.0:
            return replace(self)
        return replace(self,
                       gradient_norm_threshold=self.gradient_norm_threshold * scale,
                       cost_decrease_threshold=self.cost_decrease_threshold * scale,
                       parameter_variation_threshold=self.parameter_variation_threshold * scale,
                       irls_cost_difference_threshold=self.irls_cost_difference_threshold * scale)


def solve_rounds(round_solver, initial_estimate, options=None, regularization_parameter_sum=0.0):
    """The round structure of IRLSMapSolver::Solve (irls_map_solver.cpp:192-265) around
    round_solver(c0, c1, x0_slice, scaled_options) -> (x_slice, report): one round over all channels,
    or one per channel with split_channels; thresholds scaled for the round's parameter count."""
    opt = options if options is not None else IrlsMapSolverOptions()
    x0 = np.ascontiguousarray(initial_estimate, dtype=np.float64)
    Cn, H, W = x0.shape
    per_split = 1 if opt.split_channels else Cn               # :200-206
    rounds = Cn // per_split
    scaled = opt.adjusted(per_split * H * W, regularization_parameter_sum)   # :213-216
    out = np.empty_like(x0)
    reports = []
    for i in range(rounds):
        c0, c1 = i * per_split, (i + 1) * per_split
        x, rep = round_solver(c0, c1, x0[c0:c1], scaled)
        out[c0:c1] = x
        reports.append(rep)
    return out, reports


def solve(engine, initial_estimate, options=None, regularization_parameter_sum=0.0):
    """IRLSMapSolver::Solve on `engine` (model, observations and regularizer already set).
    initial_estimate: [C][H][W].  Returns (estimate [C][H][W], list of per-round report dicts)."""
    x0 = np.ascontiguousarray(initial_estimate, dtype=np.float64)
    assert x0.shape == (engine.C, engine.H, engine.W), x0.shape

    def device_round(c0, c1, x0_slice, scaled):
        engine.set_channel_range(c0, c1)
        return engine.solve_irls(x0_slice, epsg=scaled.gradient_norm_threshold,
                                 epsf=scaled.cost_decrease_threshold,
                                 epsx=scaled.parameter_variation_threshold,
                                 maxits=scaled.max_num_solver_iterations,
                                 max_irls_iterations=scaled.max_num_irls_iterations,
                                 irls_cost_difference_threshold=scaled.irls_cost_difference_threshold,
                                 lbfgs_corrections=(scaled.num_lbfgs_hessian_corrections
                                                    if scaled.least_squares_solver == "lbfgs" else 0))
    try:
        return solve_rounds(device_round, x0, options, regularization_parameter_sum)
    finally:
        engine.set_channel_range(0, engine.C)
