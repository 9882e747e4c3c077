import sys, os, time
sys.path.insert(0, os.getcwd())
from cerberus_b200 import abi, synth, lib
cfg = abi.default_config(); cfg.max_batch, cfg.max_features, cfg.max_obs = 1024, 160, 160 * 11
gb = lib.Backend(cfg)
batch = synth.generate_batch(8, 150, gb, prior_features=24)
big = synth.tile_batch(batch, 1024); saved = big.copy_states()
for it in range(5):
    big.restore_states(saved)
    t0 = time.perf_counter(); gb.solve_batch(big); print("e2e ms", (time.perf_counter() - t0) * 1e3)
