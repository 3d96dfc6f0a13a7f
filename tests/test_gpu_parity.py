"""GPU parity tests proper: the sm_100a kernels, called through the C ABI (libmocca_b200.so via ctypes), against
the float64 CPU oracle on identical seeded inputs.  Oracle = our restatement of Bullet's pipeline (PyBullet golden
vectors unavailable: "vs restatement").  Tolerances follow BASELINE.json's north_star."""
import numpy as np
import pytest

from tests.helpers import contact_states, oracle_state, random_states, state_error

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _env(n, seed=0, **kw):
    from mocca_envs_b200.vec_env import Walker3DCustomVecEnv

    return Walker3DCustomVecEnv(n, device="cuda:0", seed=seed, **kw)


def test_mass_matrix_and_inverse_dynamics(walker_table, oracle_mod, torch_mod):
    """north_star: mass matrix and inverse dynamics within 1e-4 relative error."""
    torch, O, t = torch_mod, oracle_mod, walker_table
    A = t["n_dof"]
    m = O.model_from_table(t)
    rng = np.random.RandomState(0)
    N = 64
    st = random_states(t, rng, N)
    env = _env(N)
    env.set_state(torch.tensor(st, dtype=torch.float32))
    M = env.mass_matrix().cpu().numpy()
    acc = rng.randn(N, 6 + A)
    tau = env.inverse_dynamics(torch.tensor(acc, dtype=torch.float32)).cpu().numpy()
    for i in range(N):
        s = oracle_state(O, A, st[i].astype(np.float32).astype(np.float64))
        Mref = O.mass_matrix(m, s)
        assert np.abs(M[i] - Mref).max() / np.abs(Mref).max() < 1e-4
        assert np.allclose(M[i], M[i].T)
        ref = O.rnea(m, s, acc[i].astype(np.float32).astype(np.float64), 9.8)
        assert np.abs(tau[i] - ref).max() / np.abs(ref).max() < 1e-4
    env.close()


def test_contact_free_single_step(walker_table, oracle_mod, torch_mod):
    """north_star: contact-free single-step state within 1e-4."""
    torch, O, t = torch_mod, oracle_mod, walker_table
    A = t["n_dof"]
    m = O.model_from_table(t)
    p = O.default_params()
    rng = np.random.RandomState(1)
    N = 64
    st = random_states(t, rng, N, spin=0.5, margin=0.35).astype(np.float32)
    tau = (np.array(t["gain"]) * rng.uniform(-1, 1, (N, A))).astype(np.float32)
    env = _env(N)
    env.set_state(torch.tensor(st))
    rows, nc = env.step_physics(torch.tensor(tau))
    out = env.get_state().cpu().numpy()
    assert int(rows.sum()) == 0 and int(nc.sum()) == 0
    worst = 0.0
    for i in range(N):
        s = oracle_state(O, A, st[i].astype(np.float64))
        O.step_physics(m, p, s, tau[i].astype(np.float64))
        worst = max(worst, state_error(out[i], O.state_vector(s, A)))
    assert worst < 1e-4, worst
    env.close()


def test_contact_single_frame(walker_table, oracle_mod, torch_mod):
    """north_star: per-step contact rollouts within a stated tolerance over 1 frame.
    Stated tolerance: 2e-3 of max(1, |x|) per state component (f32 PGS in the factor-transformed space vs f64
    velocity-space PGS), identical contact counts."""
    torch, O, t = torch_mod, oracle_mod, walker_table
    A = t["n_dof"]
    m = O.model_from_table(t)
    p = O.default_params()
    rng = np.random.RandomState(2)
    N = 32
    st = contact_states(O, t, rng, N).astype(np.float32)
    tau = (0.3 * np.array(t["gain"]) * rng.uniform(-1, 1, (N, A))).astype(np.float32)
    env = _env(N)
    env.set_state(torch.tensor(st))
    rows, nc = env.step_physics(torch.tensor(tau))
    out = env.get_state().cpu().numpy()
    rows, nc = rows.cpu().numpy(), nc.cpu().numpy()
    worst = 0.0
    for i in range(N):
        s = oracle_state(O, A, st[i].astype(np.float64))
        c, r = O.step_physics(m, p, s, tau[i].astype(np.float64))
        assert c.n == nc[i]
        assert abs(r - rows[i]) <= 2
        worst = max(worst, state_error(out[i], O.state_vector(s, A)))
    assert worst < 2e-3, worst
    env.close()


def test_self_contact_single_frame(walker_table, oracle_mod, torch_mod):
    """SURVEY 8 f1, self-collision (robots.py:259-264): limb-vs-limb contacts, each row coupling two links of the
    multibody.  Stated tolerance: >= 90% of the states within the ground-contact bound (2e-3), median < 2e-4, all
    within 2e-2, identical contact counts; with physics={"self_collision": 0} the same states must disagree (the
    feature is live on the device)."""
    from tests.helpers import self_contact_states

    torch, O, t = torch_mod, oracle_mod, walker_table
    A = t["n_dof"]
    m = O.model_from_table(t)
    p = O.default_params()
    rng = np.random.RandomState(3)
    N = 32
    st = self_contact_states(O, t, rng, N).astype(np.float32)
    tau = (np.array(t["gain"]) * rng.uniform(-1, 1, (N, A))).astype(np.float32)
    refs, counts, rws = [], [], []
    for i in range(N):
        s = oracle_state(O, A, st[i].astype(np.float64))
        c, r = O.step_physics(m, p, s, tau[i].astype(np.float64))
        refs.append(O.state_vector(s, A))
        counts.append(c.n)
        rws.append(r)
    worst = {}
    for flag in (1, 0):
        env = _env(N, physics={"self_collision": flag})
        env.set_state(torch.tensor(st))
        rows, nc = env.step_physics(torch.tensor(tau))
        out = env.get_state().cpu().numpy()
        rows, nc = rows.cpu().numpy(), nc.cpu().numpy()
        if flag:
            assert list(nc) == counts
            assert all(abs(int(a) - int(b)) <= 2 for a, b in zip(rows, rws))
        errs = np.array([state_error(out[i], refs[i]) for i in range(N)])
        worst[flag] = errs.max()
        if flag:
            # a hand hitting the pelvis at 80 rad/s is a stiff event (ERP/dt = 216 1/s on a 0.2 kg link): the oracle
            # itself moves by 1e-4 under 1e-7 relative input noise there, so a few states exceed the 2e-3 bound
            assert (errs < 2e-3).mean() >= 0.9, errs
            assert np.median(errs) < 2e-4, errs
        env.close()
    assert worst[1] < 2e-2, worst
    assert worst[0] > 0.1, worst


@pytest.mark.parametrize("eval_mode", [0, 1])
def test_reset_bit_exact_vs_numpy(eval_mode, walker_table, oracle_mod, torch_mod):
    """north_star: bit-exact reset-state generation from the same seed (values rounded to the f32 state); in
    evaluation mode randomize_target consumes one word of the env stream instead of five (env_locomotion.py:67-74)."""
    torch, O, t = torch_mod, oracle_mod, walker_table
    N = 16
    env = _env(N, seed=100)
    oracles = [O.Walker3DCustomOracle(t, seed=100 + i) for i in range(N)]
    if eval_mode:
        env.evaluation_mode()
        for o in oracles:
            o.e.eval_mode = 1
    for episode in range(3):
        obs = env.reset().cpu().numpy()
        st = env.get_state().cpu().numpy()
        rec = env.get_record().cpu().numpy()
        for i, o in enumerate(oracles):
            oref = o.reset()
            assert np.array_equal(st[i, 13:34], np.array(o.e.s.q[:21]).astype(np.float32))
            assert np.array_equal(st[i, 0:7], np.array([0, 0, 1.32, 0, 0, 0, 1], dtype=np.float32))
            assert np.array_equal(rec[i, 0:3], np.array(o.e.walk_target[:], dtype=np.float32))
            assert rec[i, 5] == o.e.stop_frames
            assert np.abs(obs[i] - oref).max() < 1e-5
    env.close()


def test_env_step_teacher_forced(oracle_mod, torch_mod):
    """Walker3DCustomEnv.step obs / reward / done on the device from f32-identical states and bookkeeping, 16 envs x 40
    steps, random actions of amplitude 1.2.  Steps outside 1e-3 (obs) / 1e-2 (reward) must be explained by a verified
    discontinuity and bounded (tests/teacher.py), else the test fails; integer bookkeeping read back and compared
    exactly after every structurally identical step."""
    from tests import teacher as T

    js = T.run_vs_oracle(oracle_mod, "walker3d", "gpu", range(7, 23), 40, lambda rng, k: rng.uniform(-1.2, 1.2, 21))
    assert np.median(np.concatenate([j.errs for j in js])) < 2e-4


def test_rollout_statistics_vs_oracle(walker_table, oracle_mod, torch_mod):
    """north_star: episode return and length for fixed random policies statistically indistinguishable.
    Stated bound: |mean_gpu - mean_oracle| < 4 standard errors (pooled) for episode length and return,
    random-uniform policy, >= 256 oracle episodes vs >= 4096 GPU episodes."""
    torch, O, t = torch_mod, oracle_mod, walker_table
    N = 2048
    env = _env(N, seed=11)
    env.reset()
    g = torch.Generator(device="cuda:0").manual_seed(5)
    lens, rets = [], []
    for _ in range(120):
        a = torch.rand(N, 21, device="cuda:0", generator=g) * 2 - 1
        obs, rew, done, info = env.step(a)
        d = done.bool()
        if d.any():
            rec = env.get_record()
            lens.append(rec[d, 20].view(torch.int32).float().cpu().numpy())
            rets.append(rec[d, 19].cpu().numpy())
    lens, rets = np.concatenate(lens), np.concatenate(rets)
    assert len(lens) >= 4096
    olens, orets = [], []
    arng = np.random.RandomState(9)
    o = O.Walker3DCustomOracle(t, seed=12345)
    while len(olens) < 256:
        o.reset()
        L, R = 0, 0.0
        while True:
            _, r, d, _ = o.step(arng.uniform(-1, 1, 21))
            L += 1
            R += r
            if d:
                break
        olens.append(L)
        orets.append(R)
    olens, orets = np.array(olens), np.array(orets)
    for a_, b_ in ((lens, olens), (rets, orets)):
        se = np.sqrt(a_.var() / len(a_) + b_.var() / len(b_))
        assert abs(a_.mean() - b_.mean()) < 4 * se + 1e-6, (a_.mean(), b_.mean(), se)
    env.close()


def test_full_size_properties(torch_mod):
    """BASELINE config 2 size (16384 envs): determinism, finiteness, auto-reset bookkeeping, host path == device."""
    torch = torch_mod
    N = 16384
    outs = []
    for rep in range(2):
        env = _env(N, seed=3)
        env.reset()
        g = torch.Generator(device="cuda:0").manual_seed(1)
        tot_done = 0
        for _ in range(40):
            a = torch.rand(N, 21, device="cuda:0", generator=g) * 2 - 1
            obs, rew, done, info = env.step(a)
            tot_done += int(done.sum())
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
        st = env.stats()
        assert st["episodes"] == tot_done and st["nonfinite"] == 0
        assert tot_done > 0
        outs.append((obs.clone(), env.get_state().clone()))
        env.close()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_batch_size_independence(torch_mod):
    """Env i evolves identically whatever the batch size (CTA padding / tail warps must have no side effects):
    N = 5, 13 and 24 with the same per-env seeds and actions."""
    torch = torch_mod
    rng = np.random.RandomState(4)
    acts = rng.uniform(-1, 1, (30, 24, 21)).astype(np.float32)
    results = {}
    for N in (5, 13, 24):
        env = _env(N, seed=50)
        env.reset()
        done_total = 0
        for k in range(30):
            obs, rew, done, _ = env.step(torch.tensor(acts[k, :N]))
            done_total += int(done.sum())
        results[N] = (obs.cpu().numpy().copy(), env.get_state().cpu().numpy().copy(), done_total, env.stats()["episodes"])
        env.close()
    for N in (13, 24):
        assert np.array_equal(results[5][0], results[N][0][:5])
        assert np.array_equal(results[5][1], results[N][1][:5])
    for N in (5, 13, 24):
        assert results[N][2] == results[N][3]  # pad envs never leak into the statistics


def test_step_host_matches_device(torch_mod):
    torch = torch_mod
    N = 256
    e1, e2 = _env(N, seed=21), _env(N, seed=21)
    e1.reset(); e2.reset()
    rng = np.random.RandomState(0)
    for _ in range(5):
        a = rng.uniform(-1, 1, (N, 21)).astype(np.float32)
        o1, r1, d1, _ = e1.step(torch.tensor(a))
        o2, r2, d2, t2 = e2.step_host(a)
        assert np.array_equal(o1.cpu().numpy(), o2) and np.array_equal(r1.cpu().numpy(), r2)
        assert np.array_equal(d1.cpu().numpy(), d2)
    e1.close(); e2.close()


def test_step_host_pinned_zero_copy_matches_device(torch_mod):
    """Pinned host buffers take the zero-copy route of mb200_step_host (the step kernel reads the actions from and
    stores obs / reward / done / trunc into the mapped host buffers itself); pageable buffers take the staged copies.
    Both must equal the device-buffer step bit for bit, through auto-resets and with pad envs in the last CTA."""
    torch = torch_mod
    N = 1000
    e1, e2 = _env(N, seed=33), _env(N, seed=33)
    e1.reset(); e2.reset()
    h_act = torch.empty(N, 21).pin_memory()
    outs = (torch.empty(N, 52).pin_memory(), torch.empty(N).pin_memory(),
            torch.empty(N, dtype=torch.uint8).pin_memory(), torch.empty(N, dtype=torch.uint8).pin_memory())
    outs_np = tuple(o.numpy() for o in outs)
    rng = np.random.RandomState(1)
    dones = 0
    for _ in range(60):
        a = rng.uniform(-1, 1, (N, 21)).astype(np.float32)
        h_act.numpy()[:] = a
        o1, r1, d1, i1 = e1.step(torch.tensor(a))
        for o in outs_np:
            o.fill(0)
        e2.step_host(h_act.numpy(), outs_np)
        assert np.array_equal(o1.cpu().numpy(), outs_np[0]) and np.array_equal(r1.cpu().numpy(), outs_np[1])
        assert np.array_equal(d1.cpu().numpy(), outs_np[2])
        assert np.array_equal(i1["TimeLimit.truncated"].cpu().numpy().astype(np.uint8), outs_np[3])
        dones += int(outs_np[2].sum())
    assert dones > N // 2  # random actions: most walkers fell and were reset at least once
    e1.close(); e2.close()


def test_gym_facade(torch_mod):
    from mocca_envs_b200 import make

    env = make("mocca_envs:Walker3DCustomEnv-v0", seed=0)
    obs = env.reset()
    assert obs.shape == (52,) and obs.dtype == np.float64
    total, steps = 0.0, 0
    done = False
    while not done and steps < 1000:
        obs, r, done, info = env.step(np.zeros(21))
        total += r
        steps += 1
    assert done and steps < 200  # the passive walker collapses (oracle: 33 steps with seed 0)
    assert env.reset().shape == (52,)
    env.close()


def test_abi_errors(torch_mod):
    import ctypes as C

    from mocca_envs_b200 import _lib

    L = _lib.lib()
    h = C.c_void_p()
    assert L.mb200_create(b"Nope-v0", 4, 0, None, C.byref(h)) != 0
    assert L.mb200_create(b"Walker3DCustomEnv-v0", 0, 0, None, C.byref(h)) != 0
    assert L.mb200_create(b"Walker3DCustomEnv-v0", 4, 99, None, C.byref(h)) != 0
    assert L.mb200_step(None, None, None, None, None, None, None, None) != 0


def test_checkpoint_resume_bit_exact(torch_mod):
    """state_dict / load_state_dict (physics state, record, MT19937 streams): a restored batch -- even a freshly
    created one, whose scheduler order differs -- continues bit-exactly, through auto-resets."""
    torch = torch_mod
    N = 512
    g = torch.Generator(device="cuda:0").manual_seed(9)
    acts = torch.rand(40, N, 21, device="cuda:0", generator=g) * 2 - 1
    env = _env(N, seed=21)
    env.reset()
    for k in range(15):
        env.step(acts[k])
    ckpt = env.state_dict()
    ref = []
    for k in range(15, 40):
        obs, rew, done, info = env.step(acts[k])
        ref.append((obs.clone(), rew.clone(), done.clone()))
    assert sum(int(d.sum()) for _, _, d in ref) > 50  # resets (terrain / pose draws from the streams) were exercised
    env.close()
    env2 = _env(N, seed=999)  # different seed: everything must come from the checkpoint
    env2.reset()
    env2.load_state_dict(ckpt)
    for k in range(15, 40):
        obs, rew, done, info = env2.step(acts[k])
        o, r, d = ref[k - 15]
        assert torch.equal(obs, o) and torch.equal(rew, r) and torch.equal(done, d), k
    env2.close()


def test_warm_start_switch(walker_table, oracle_mod, torch_mod):
    """Bullet-version switch mb200_physics.warmstart on the device: persistent per-candidate
    contact impulses in HBM, across stepSimulation calls.  Flipping the switch moves the oracle and the device TOGETHER:
    three consecutive steps from in-contact states agree within the contact-frame tolerance (2e-3) with the switch on, the
    impulse arrays agree, and the switched-on states differ from the switched-off ones by far more than the tolerance."""
    import ctypes as C

    torch, O, t = torch_mod, oracle_mod, walker_table
    A = 21
    m = O.model_from_table(t)
    rng = np.random.RandomState(5)
    N = 32
    st = contact_states(O, t, rng, N).astype(np.float32)
    tau = (0.3 * np.array(t["gain"]) * rng.uniform(-1, 1, (N, A))).astype(np.float32)
    outs = {}
    for f in (0.0, 0.85):
        env = _env(N, physics={"warmstart": f})
        env.set_state(torch.tensor(st))
        for k in range(3):
            env.step_physics(torch.tensor(tau))
        outs[f] = env.get_state().cpu().numpy()
        if f > 0:
            warm_dev = env.state_dict()["warm"].numpy()
            p = O.default_params()
            p.warmstart = f
            worst = 0.0
            for i in range(N):
                s = oracle_state(O, A, st[i].astype(np.float64))
                warm = (C.c_double * 384)()
                for k in range(3):
                    O.step_physics(m, p, s, tau[i].astype(np.float64), warm=warm)
                worst = max(worst, state_error(outs[f][i], O.state_vector(s, A)))
                wo = np.array(warm[:])
                assert np.abs(warm_dev[i] - wo).max() < 2e-3 * max(1.0, np.abs(wo).max()), i
            assert worst < 2e-3, worst
        env.close()
    assert max(state_error(outs[0.85][i], outs[0.0][i]) for i in range(N)) > 2e-2
