"""Env-worker feed: the learner's side of the per-step exchange with the env worker processes
(train.py:615-654, envs.py:305-340), without the float64 round trip.

The reference receives, per env and per step, a pickled ``[state float64 [4,84,84], reward, done, trunc, visited_rooms]``
message over a ``multiprocessing.Pipe`` (225 KB each), copies it into float64 numpy batches and converts those to float32
tensors for ``get_action`` / the rollout buffer.  The frames are raw ALE pixels: integers 0..255.  Two pieces here:

* ``StepCollector`` speaks the reference's pipe protocol UNCHANGED (same message order, including the episode-end follow-up
  messages, so the reference's env workers run as they are) but assembles the step directly into a pinned uint8 staging
  batch and returns it as device tensors: one 28 KB-per-env H2D copy, ready for ``RNDAgent.get_action`` (uint8 frames are
  divided by 255 in the patch kernel, bit-identical to ``np.float32(x) / 255.``) and ``DeviceRollout.add``.
* ``FrameRing`` is a shared-memory uint8 ring for workers that can be changed: a worker writes its frame stack straight into
  its slot and sends only the scalars over the pipe (8x fewer bytes than float64, no pickling of the frames).

The environments themselves (ALE) are not part of this package.
"""
from __future__ import annotations

from multiprocessing import shared_memory
from typing import List, Optional, Sequence

import numpy as np
import torch


class FrameRing:
    """Shared-memory frame slots ``[slots][num_env][stack][H][W]`` uint8.  The learner creates it, workers attach by name."""

    def __init__(self, num_env: int, stack: int = 4, image: int = 84, slots: int = 2, name: Optional[str] = None):
        self.shape = (slots, num_env, stack, image, image)
        nbytes = int(np.prod(self.shape))
        self.owner = name is None
        self.shm = shared_memory.SharedMemory(create=True, size=nbytes) if name is None else shared_memory.SharedMemory(name=name)
        self.frames = np.ndarray(self.shape, dtype=np.uint8, buffer=self.shm.buf)
        self.name = self.shm.name

    @classmethod
    def attach(cls, name: str, num_env: int, stack: int = 4, image: int = 84, slots: int = 2) -> "FrameRing":
        return cls(num_env, stack, image, slots, name=name)

    def write(self, slot: int, env_idx: int, state) -> None:
        """Worker side: ``state`` [stack,H,W] with pixel values 0..255 (uint8, or the reference's float array)."""
        self.frames[slot % self.shape[0], env_idx] = state          # numpy casts float -> uint8 exactly for integral values

    def batch(self, slot: int) -> np.ndarray:
        return self.frames[slot % self.shape[0]]

    def close(self) -> None:
        self.frames = None
        self.shm.close()
        if self.owner:
            self.shm.unlink()


class StepCollector:
    """train.py:615-654 for one rank's env workers.

    ``step(actions)`` sends one action per worker and gathers the replies in the reference's order.  Returns a dict:
    ``states`` uint8 [E,stack,H,W] and ``next_obs`` uint8 [E,1,H,W] (the newest frame, train.py:640) as tensors on
    ``device`` (host tensors when ``device`` is None), ``rewards`` float64 [E], ``dones`` bool [E], ``truncs`` bool [E]
    (numpy), ``visited_rooms`` (union of the sets sent this step) and ``episodes`` -- one dict per finished episode with
    the follow-up fields of envs.py:333-336 (``env_idx``, ``undiscounted_episode_return``, ``l``,
    ``num_finished_episodes`` and, for Montezuma, ``number_of_visited_rooms`` / ``visited_rooms``)."""

    def __init__(self, parent_conns: Sequence, stack: int = 4, image: int = 84, device=None, montezuma: bool = False,
                 ring: Optional[FrameRing] = None):
        self.conns = list(parent_conns)
        self.E, self.stack, self.image = len(self.conns), stack, image
        self.device = torch.device(device) if device is not None else None
        self.montezuma = montezuma
        self.ring = ring
        self._slot = 0
        # Two pinned staging batches used alternately, each guarded by the CUDA event of its last upload: the host never
        # rewrites a buffer whose asynchronous H2D copy may still be in flight (a device-side consumer such as
        # DeviceRollout.add does not force a sync between steps the way get_action(...).cpu() does).
        cuda = self.device is not None and self.device.type == "cuda"
        self._stages = [torch.empty(self.E, stack, image, image, dtype=torch.uint8) for _ in range(2 if cuda else 1)]
        if cuda:
            self._stages = [t.pin_memory() for t in self._stages]
        self._copied = [None] * len(self._stages)
        self._cur = 0
        self._stage = self._stages[0]
        self._stage_np = self._stage.numpy()

    def initial_states(self):
        """envs.py:305: every worker first sends its reset state."""
        for i, c in enumerate(self.conns):
            self._stage_np[i] = c.recv()
        return self._to_device()

    def _to_device(self):
        if self.device is None:
            st = self._stage.clone()
        else:
            st = self._stage.to(self.device, non_blocking=True)
            if self.device.type == "cuda":
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                self._copied[self._cur] = ev
                self._cur ^= 1                               # the next step fills the other buffer ...
                if self._copied[self._cur] is not None:
                    self._copied[self._cur].synchronize()    # ... once its own upload (two steps ago) has completed
                self._stage = self._stages[self._cur]
                self._stage_np = self._stage.numpy()
        return st, st[:, self.stack - 1:self.stack]

    def step(self, actions) -> dict:
        for c, a in zip(self.conns, actions):
            c.send(a)
        rewards = np.zeros(self.E, dtype=np.float64)
        dones = np.zeros(self.E, dtype=np.bool_)
        truncs = np.zeros(self.E, dtype=np.bool_)
        rooms: set = set()
        episodes: List[dict] = []
        slot = self._slot
        for i, c in enumerate(self.conns):
            s, r, d, trun, visited = c.recv()
            if s is None:                                    # ring worker: the frames are already in shared memory
                self._stage_np[i] = self.ring.batch(slot)[i]
            else:
                self._stage_np[i] = s                        # float64 pixels 0..255 -> uint8, exact
            rewards[i], dones[i], truncs[i] = r, d, trun
            if visited:
                rooms |= set(visited)
            if d or trun:                                    # envs.py:333-336: follow-up messages of a finished episode
                ep = {"env_idx": i}
                if self.montezuma:
                    ep["number_of_visited_rooms"], ep["visited_rooms"] = c.recv()
                ep["undiscounted_episode_return"], ep["l"], ep["num_finished_episodes"] = c.recv()
                episodes.append(ep)
        self._slot += 1
        states, next_obs = self._to_device()
        return dict(states=states, next_obs=next_obs, rewards=rewards, dones=dones, truncs=truncs, visited_rooms=rooms,
                    episodes=episodes)
