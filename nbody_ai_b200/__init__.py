"""B200-native batched N-body TTV engine: drop-in for the generate()/analyze() path of
pearsonkyle/Nbody-AI.  See DESIGN.md and INTEGRATION.md.

    nbody_ai_b200.simulation   generate, integrate, analyze, transit_times, TTV, randomize (+ *_batch)
    nbody_ai_b200.tools        constants, sa, rv_semi, hill_sphere, maxavg, lomb_scargle, find_zero, TTV
    nbody_ai_b200.nested       get_ttv, get_ttv_batch, loglike_batch (nbody/nested.py forward model)
    nbody_ai_b200.ttv_fit      NbodyLikelihood (ttv_fit.py likelihood, vectorised)
    nbody_ai_b200.grid         generate_sim, simulation_grid (simulations/simulation_grid.py)
    nbody_ai_b200.dataset      generate_training_set (simulations/generate_simulations.py)
    nbody_ai_b200.engine       Plan / Buffers / run / run_device / run_multi_gpu over the C ABI (include/nbt.h)
"""
