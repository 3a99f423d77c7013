"""Batched counterpart of the reference's training mission: the cadence of
``missions/coma_mission.py:COMAMission.execute`` — collect, update, log scalars, keep the best model, evaluate —
on top of ``COMATrainer`` (8192 environments per rollout instead of one).

Kept from the reference: the scalar tags of ``add_to_tensorboard`` (coma_mission.py:174-423: ``{mode}Return/Episode/
mean|std|max|min`` over the absolute returns, ``{mode}Rewards/Episode/...`` over the per-step rewards,
``{mode}Return/Relative(used)/Episode/...`` over the relative returns that train the critic), the best-model rule of
``save_best_model`` (:425-435: running mean over all updates so far of the return collected before each update,
compared once ``patience`` updates exist; the actor is saved to ``best_model.pth``), the evaluation cadence (:123-170:
every ``eval_every`` training steps, greedy episodes, same scalars with mode ``eval``) and the per-step evaluation
curves of the baselines (IG_baseline.py:81-100,191-210: masked entropy and F1 of the global map) from
``ipp_eval_metrics``.  Not reproduced: the matplotlib figures (action / altitude bar plots, trajectories) — their data
is logged as scalars (``{mode}Actions/a``, ``{mode}Altitudes/z``) instead.

One *update* here = one rollout of every env of the batch + ``data_passes`` passes over it; the reference fires an
update every 5 single-env episodes (300 transitions).
"""
import json
import os

import torch


class ScalarLog:
    """``add_scalar(tag, value, step)`` sink: a TensorBoard ``SummaryWriter`` when given (or importable and
    ``tensorboard=True``), and always a JSON-lines file ``scalars.jsonl`` in ``log_dir``."""

    def __init__(self, log_dir, writer=None, tensorboard=False):
        os.makedirs(log_dir, exist_ok=True)
        self.log_dir = log_dir
        self.writer = writer
        if writer is None and tensorboard:
            from torch.utils.tensorboard import SummaryWriter  # the reference's writer (coma_mission.py:9)

            self.writer = SummaryWriter(log_dir)
        self._f = open(os.path.join(log_dir, "scalars.jsonl"), "a")
        self.last = {}

    def add_scalar(self, tag, value, step):
        value = float(value)
        self.last[tag] = value
        self._f.write(json.dumps({"tag": tag, "value": value, "step": int(step)}) + "\n")
        if self.writer is not None:
            self.writer.add_scalar(tag, value, step)

    def flush(self):
        self._f.flush()
        if self.writer is not None:
            self.writer.flush()

    def close(self):
        self.flush()
        self._f.close()


class BestModel:
    """coma_mission.py:425-435: ``running_mean = mean(returns of all updates so far)``; once ``patience`` updates exist
    and the running mean exceeds the best seen, it becomes the best and the actor is saved."""

    def __init__(self, patience, path, best=float("-inf")):
        self.patience = int(patience)
        self.path = path
        self.returns = []
        self.best = best

    def offer(self, episode_return, actor):
        self.returns.append(float(episode_return))
        running = sum(self.returns) / len(self.returns)
        if len(self.returns) >= self.patience and running > self.best:
            self.best = running
            if actor is not None and self.path is not None and _rank() == 0:  # every rank keeps the books, one writes
                save_actor(actor, self.path)
            return True
        return False


def _rank():
    import torch.distributed as dist

    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def save_actor(actor, path):
    """Checkpoint format: the actor's ``state_dict()`` (tensors only).  The reference pickles the whole
    ``ActorNetwork`` module (missions/coma_mission.py:435 ``torch.save(actor_network, ...)``), which ties a checkpoint
    to the reference's class path; ``load_actor_state`` below reads both forms."""
    torch.save(actor.state_dict(), path)


def load_actor_state(path, map_location="cpu"):
    """State dict of an actor checkpoint written either by ``save_actor`` (a state dict) or by the reference
    (a pickled ``ActorNetwork`` module, coma_mission.py:435 / coma_test.py:52 — needs the reference's classes
    importable, e.g. after ``facade.install``).  Layer names are the same in both (conv1-3, fc1-3)."""
    try:
        obj = torch.load(path, map_location=map_location, weights_only=True)
    except Exception:
        obj = torch.load(path, map_location=map_location, weights_only=False)
    if hasattr(obj, "state_dict"):
        obj = obj.state_dict()
    return {k: v for k, v in obj.items()}


def next_episode_ids(n_envs, next_episode, world):
    """Episode numbers of this rank's next rollout and the rank's counter after it: every rollout of the job consumes
    ``n_envs * world`` consecutive numbers, rank r's block starts at ``env_id_base + 1 = r * n_envs + 1``."""
    ids = torch.arange(int(n_envs), dtype=torch.int64) + int(next_episode)
    return ids, int(next_episode) + int(n_envs) * int(world)


def _stats(log, prefix, values, step):
    v = values.double().flatten()
    log.add_scalar(prefix + "/mean", v.mean(), step)
    log.add_scalar(prefix + "/std", v.std(unbiased=False), step)  # np.std
    log.add_scalar(prefix + "/max", v.max(), step)
    log.add_scalar(prefix + "/min", v.min(), step)


class COMAMission:
    def __init__(self, trainer, log_dir, eval_every=50, eval_rollouts=1, writer=None, tensorboard=False,
                 max_mean_episode_return=float("-inf")):
        self.trainer = trainer
        self.env = trainer.env
        self.log = ScalarLog(log_dir, writer=writer, tensorboard=tensorboard)
        patience = trainer.params["experiment"]["missions"]["patience"]
        self.best = BestModel(patience, os.path.join(log_dir, "best_model.pth"), best=max_mean_episode_return)
        self.eval_every = int(eval_every)
        self.eval_rollouts = int(eval_rollouts)
        self.training_step_idx = 0
        self.environment_step_idx = 0
        self._next_episode = self.env.env_id_base + 1

    # ---- scalars of one batch of episodes (coma_mission.py:174-260) ------------------------------------------
    def _log_episodes(self, mode, step):
        tr, env = self.trainer, self.env
        rel = tr.buf_rew  # [T, B] relative rewards (the ones the critic is trained on)
        ab = tr.buf_abs   # [T, B] absolute rewards
        _stats(self.log, "%sReturn/Episode" % mode, ab.sum(0), step)
        _stats(self.log, "%sRewards/Episode" % mode, rel, step)
        _stats(self.log, "%sReturn/Relative(used)/Episode" % mode, rel.sum(0), step)
        acts = tr.buf_act.flatten()
        for a in range(6):
            self.log.add_scalar("%sActions/%d" % (mode, a), (acts == a).sum(), step)
        alt = env.positions[1:env.T + 1, ..., 2].flatten()  # altitudes flown to
        for z in range(env.tables.min_altitude, env.tables.max_altitude + 1, env.tables.spacing):
            self.log.add_scalar("%sAltitudes/%d" % (mode, z), (alt == z).sum(), step)

    def _episodes(self):
        ids, self._next_episode = next_episode_ids(self.env.B, self._next_episode, self._world())
        return ids

    @staticmethod
    def _world():
        import torch.distributed as dist

        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    # ---- evaluation: greedy episodes + the baselines' metric curves -------------------------------------------
    @torch.no_grad()
    def evaluate(self):
        tr, env = self.trainer, self.env
        done_before, eps_before = tr.episodes_done, tr._last_eps_episode
        ent_curve = f1_curve = None
        for _ in range(self.eval_rollouts):
            tr.rollout(episodes=self._episodes(), greedy=True)
            self._log_episodes("eval", self.training_step_idx)
        # masked entropy / F1 of the accumulated global map after the last episode (IG_baseline.py:191-210)
        env.observe(final=True)
        ent, f1 = env.eval_metrics()
        ent_curve, f1_curve = ent.mean(), f1.mean()
        self.log.add_scalar("evalMetrics/entropy_final", ent_curve, self.training_step_idx)
        self.log.add_scalar("evalMetrics/f1_final", f1_curve, self.training_step_idx)
        tr.episodes_done, tr._last_eps_episode = done_before, eps_before  # evaluation does not advance the eps schedule
        return float(ent_curve), float(f1_curve)

    # ---- the training loop (coma_mission.py:49-172) -------------------------------------------------------------
    def execute(self, n_updates):
        tr, env = self.trainer, self.env
        for _ in range(int(n_updates)):
            mean_return = tr.rollout(episodes=self._episodes())
            if self._world() > 1:  # every rank keeps the same best-model bookkeeping
                import torch.distributed as dist

                dist.all_reduce(mean_return)
                mean_return = mean_return / self._world()
            stats = tr.update()
            self.training_step_idx += 1
            self.environment_step_idx += env.T * env.A * env.B * self._world()
            step = self.training_step_idx
            self._log_episodes("train", step)
            for k, v in stats.items():
                self.log.add_scalar("Training/%s" % k, v, step)
            self.log.add_scalar("Training/environment_steps", self.environment_step_idx, step)
            self.best.offer(mean_return, tr.actor)
            if self.eval_every > 0 and step % self.eval_every == 0:
                self.evaluate()
            self.log.flush()
        return self.best.best
