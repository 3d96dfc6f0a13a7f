#!/usr/bin/env python
"""A/B of the three host-buffer result paths of mb200_step_host (MB200_HOST_DIRECT = 0 staged copies, 1 zero-copy
result stores from the step kernel, 2 also zero-copy action loads): wall time per step and a checksum of everything
the host received (the three runs must agree bit for bit)."""
import os, subprocess, sys, time, zlib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def child(N, K):
    import numpy as np
    import torch
    from mocca_envs_b200.vec_env import Walker3DCustomVecEnv
    env = Walker3DCustomVecEnv(N, device="cuda:0", seed=1)
    env.reset()
    A, OB = env.act_dim, env.obs_dim
    g = torch.Generator().manual_seed(5)
    pool = (torch.rand(16, N, A, generator=g) * 2 - 1).pin_memory()
    h_obs = torch.empty(N, OB).pin_memory(); h_rew = torch.empty(N).pin_memory()
    h_done = torch.empty(N, dtype=torch.uint8).pin_memory(); h_trunc = torch.empty(N, dtype=torch.uint8).pin_memory()
    outs = (h_obs.numpy(), h_rew.numpy(), h_done.numpy(), h_trunc.numpy())
    pn = pool.numpy()
    crc = 0
    for k in range(60):
        env.step_host(pn[k % 16], outs)
        for o in outs:
            crc = zlib.crc32(o.tobytes(), crc)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(K):
        env.step_host(pn[k % 16], outs)
    wall = (time.perf_counter() - t0) / K
    # pageable buffers must still work (staged path)
    po = (np.empty((N, OB), np.float32), np.empty(N, np.float32), np.empty(N, np.uint8), np.empty(N, np.uint8))
    env.step_host(np.array(pn[0]), po)
    assert np.isfinite(po[0]).all()
    print("MB200_HOST_DIRECT=%s N=%d: %.1f us per step = %.2f M env-steps/s, crc %08x"
          % (os.environ.get("MB200_HOST_DIRECT", "default"), N, wall * 1e6, N / wall / 1e6, crc))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2]), int(sys.argv[3]))
    else:
        N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
        for mode in ("0", "1", "2"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "child", str(N), "400"],
                                  env=dict(os.environ, MB200_HOST_DIRECT=mode))
