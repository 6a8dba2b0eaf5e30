"""Env feed: StepCollector speaks the reference's pipe protocol (envs.py:305-340 sends, train.py:615-654
receives) and must hand back exactly what the reference's receive loop assembles -- as uint8.  Worker processes here are
synthetic (seeded random frames; no ALE in this image) but send the reference's messages, follow-ups included."""
import multiprocessing as mp

import numpy as np
import pytest
import torch


def _worker(conn, seed, steps, montezuma, ring_name, env_idx, num_env):
    """envs.py:297-336 message sequence with synthetic frames."""
    rng = np.random.default_rng(seed)
    ring = None
    if ring_name is not None:
        from eavit_b200.envfeed import FrameRing
        ring = FrameRing.attach(ring_name, num_env)
    state = rng.integers(0, 256, (4, 84, 84)).astype(np.float64)
    conn.send(state)                                                   # initial reset state
    n_ep = 0
    for t in range(steps):
        action = conn.recv()
        assert isinstance(action, (int, np.integer))
        state = rng.integers(0, 256, (4, 84, 84)).astype(np.float64)
        reward = float(rng.normal())
        done = bool(rng.random() < 0.2)
        trun = bool(rng.random() < 0.1)
        rooms = {int(rng.integers(0, 24))} if montezuma else {}
        if ring is not None:
            ring.write(t, env_idx, state)
            conn.send([None, reward, done, trun, rooms])
        else:
            conn.send([state, reward, done, trun, rooms])
        if done or trun:
            n_ep += 1
            if montezuma:
                conn.send([len(rooms), rooms])
            conn.send([float(rng.normal()), int(rng.integers(1, 500)), n_ep])
    conn.close()


def _reference_receive(conns, actions, montezuma):
    """train.py:615-654, restated: float64 batches + the follow-up reads."""
    E = len(conns)
    for c, a in zip(conns, actions):
        c.send(a)
    next_states = np.zeros([E, 4, 84, 84], dtype=np.float64)
    rewards, dones = np.zeros([E], dtype=np.float64), np.zeros([E], dtype=np.bool_)
    next_obs = np.zeros([E, 1, 84, 84], dtype=np.float64)
    rooms, eps = set(), []
    for i, c in enumerate(conns):
        s, r, d, trun, visited = c.recv()
        next_states[i] = s[:]
        rewards[i], dones[i] = r, d
        rooms = rooms.union(visited)
        next_obs[i] = s[3, :, :].reshape([1, 84, 84])[:]
        if d or trun:
            ep = {"env_idx": i}
            if montezuma:
                ep["number_of_visited_rooms"], ep["visited_rooms"] = c.recv()
            ep["undiscounted_episode_return"], ep["l"], ep["num_finished_episodes"] = c.recv()
            eps.append(ep)
    return next_states, rewards, dones, next_obs, rooms, eps


def _spawn(E, steps, montezuma, ring_name=None):
    ctx = mp.get_context("spawn")
    conns, procs = [], []
    for i in range(E):
        p_conn, c_conn = ctx.Pipe()
        p = ctx.Process(target=_worker, args=(c_conn, 100 + i, steps, montezuma, ring_name, i, E), daemon=True)
        p.start()
        conns.append(p_conn)
        procs.append(p)
    return conns, procs


@pytest.mark.parametrize("montezuma", [False, True])
def test_collector_matches_reference_receive_loop(montezuma):
    import eavit_b200  # noqa
    from eavit_b200.envfeed import StepCollector
    E, steps = 3, 12
    conns_a, procs_a = _spawn(E, steps, montezuma)          # same seeds -> identical message streams
    conns_b, procs_b = _spawn(E, steps, montezuma)
    col = StepCollector(conns_a, montezuma=montezuma)
    st0, ob0 = col.initial_states()
    ref0 = np.stack([c.recv() for c in conns_b])
    assert st0.dtype == torch.uint8 and np.array_equal(st0.numpy(), ref0) and np.array_equal(ob0.numpy(), ref0[:, 3:4])
    n_eps = 0
    for t in range(steps):
        actions = [int(t % 18)] * E
        got = col.step(actions)
        ns, rw, dn, no, rooms, eps = _reference_receive(conns_b, actions, montezuma)
        assert np.array_equal(got["states"].numpy(), ns) and got["states"].dtype == torch.uint8
        assert np.array_equal(got["next_obs"].numpy(), no)
        assert np.array_equal(np.float32(got["states"].numpy()) / 255.0, np.float32(ns) / 255.0)     # what get_action is fed
        assert np.array_equal(got["rewards"], rw) and got["rewards"].dtype == np.float64
        assert np.array_equal(got["dones"], dn) and got["visited_rooms"] == rooms
        assert got["episodes"] == eps
        n_eps += len(eps)
    assert n_eps > 0                                          # the follow-up path was exercised
    for p in procs_a + procs_b:
        p.join(timeout=10)


def test_frame_ring_workers():
    """Workers write frames into shared memory and send only scalars: same batches as the pipe protocol."""
    import eavit_b200  # noqa
    from eavit_b200.envfeed import FrameRing, StepCollector
    E, steps = 3, 6
    ring = FrameRing(E)
    try:
        conns_a, procs_a = _spawn(E, steps, False, ring_name=ring.name)
        conns_b, procs_b = _spawn(E, steps, False)
        col = StepCollector(conns_a, ring=ring)
        col.initial_states()
        [c.recv() for c in conns_b]
        for t in range(steps):
            got = col.step([1] * E)
            ns, rw, dn, no, _, eps = _reference_receive(conns_b, [1] * E, False)
            assert np.array_equal(got["states"].numpy(), ns) and np.array_equal(got["rewards"], rw)
            assert [e["env_idx"] for e in got["episodes"]] == [e["env_idx"] for e in eps]
        for p in procs_a + procs_b:
            p.join(timeout=10)
    finally:
        ring.close()


@pytest.mark.gpu
def test_collector_feeds_the_device_buffer():
    """uint8 device batches go straight into get_action and DeviceRollout.add."""
    import eavit_b200  # noqa
    from eavit_b200 import rollout
    from eavit_b200.envfeed import StepCollector
    E, steps = 2, 3
    conns, procs = _spawn(E, steps, False)
    col = StepCollector(conns, device="cuda")
    st, ob = col.initial_states()
    assert st.is_cuda and st.dtype == torch.uint8 and ob.shape == (E, 1, 84, 84)
    buf = rollout.DeviceRollout(E, steps, 18)
    for t in range(steps):
        got = col.step([0] * E)
        buf.add(t, st, got["next_obs"], got["rewards"], got["dones"], np.zeros(E, np.int64), np.zeros(E, np.float32),
                np.zeros(E, np.float32), np.zeros((E, 18), np.float32), np.zeros(E, np.float32))
        assert torch.equal(buf.next_obs[:, t], got["states"][:, 3:4])
        st = got["states"]
    for p in procs:
        p.join(timeout=10)


@pytest.mark.gpu
def test_whole_loop_with_synthetic_workers():
    """train.py's loop body end to end on the drop-in pieces: StepCollector -> get_action -> compute_intrinsic_reward on
    normalised frames -> DeviceRollout -> finish -> train_model, two updates, with synthetic env worker processes."""
    import eavit_b200  # noqa
    from eavit_b200 import config, rollout, utils
    from eavit_b200.envfeed import StepCollector
    from oracle import oracle as O
    from test_gpu_model import CFGS, make_agent
    cfg = CFGS["lucid"]
    E, T, updates = 2, 4, 2
    agent, _ = make_agent(cfg, E, T)
    agent.batch_size = 4
    config.load_config(None, TrainMethod="original_RND")
    obs_rms, reward_rms = utils.RunningMeanStd(shape=(1, 1, 84, 84), usage="obs_rms"), utils.RunningMeanStd(usage="reward_rms")
    filt = utils.RewardForwardFilter(cfg.int_gamma)
    conns, procs = _spawn(E, T * updates, False)
    col = StepCollector(conns, device="cuda")
    states, obs0 = col.initial_states()
    obs_rms.update(torch.cat([obs0, 255 - obs0]))                       # some initial statistics (train.py:127-180)
    w0 = agent.state_dict()["model.actor.2.weight"].clone()
    np.random.seed(1); torch.manual_seed(1)
    for u in range(updates):
        buf = rollout.DeviceRollout(E, T, cfg.n_actions)
        for t in range(T):
            actions, v_ext, v_int, policy = agent.get_action(states)                 # uint8 frames, /255 in-kernel
            got = col.step([int(a) for a in actions])
            r_int = agent.compute_intrinsic_reward(utils.normalize_obs(got["next_obs"], obs_rms))   # train.py:666: stats from before the update
            buf.add(t, states, got["next_obs"], got["rewards"], got["dones"], actions, v_ext, v_int, policy, r_int)
            states = got["states"]
        _, v_ext, v_int, _ = agent.get_action(states)
        buf.add_last_values(v_ext, v_int)
        args = buf.finish(obs_rms, reward_rms, filt, cfg.gamma, cfg.int_gamma, cfg.lam, cfg.ext_coef, cfg.int_coef)
        agent.train_model(*args, u)
        s = agent.stats_summary()
        assert np.isfinite(s["loss"])
    torch.cuda.synchronize()
    assert not torch.equal(w0, agent.state_dict()["model.actor.2.weight"])
    assert obs_rms.count > 2 * E * T and reward_rms.count > 2 * T - 1
    for p in procs:
        p.join(timeout=10)
