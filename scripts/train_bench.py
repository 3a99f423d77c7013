#!/usr/bin/env python
"""BASELINE configs[2]/[3]: full COMA actor+critic train loop on the batched env.

    python scripts/train_bench.py --envs 8192 --iters 3                       (1 GPU)
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/train_bench.py --envs 8192   (65536 envs)

One iteration = one 15-step episode in every env (rollout: env kernels + feature kernels + actor forward) followed by
one COMA update (TD(lambda) targets, `--passes` passes of critic+actor mini-batch steps, one flattened NCCL gradient
all-reduce per network per optimizer step).  Prints one JSON line (rank 0).
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--agents", type=int, default=4)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--minibatch", type=int, default=16384)
    ap.add_argument("--bf16", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from ipp_marl_b200 import BatchedIPPEnv
    from ipp_marl_b200.coma import COMATrainer
    params = json.load(open(os.path.join(ROOT, "tests/golden/kats.json")))["synthetic50"]["params"]
    params["experiment"]["missions"]["n_agents"] = args.agents
    env = BatchedIPPEnv(params, args.envs, device=dev, env_id_base=rank * args.envs)
    tr = COMATrainer(env, params, minibatch=args.minibatch, data_passes=args.passes,
                     compute_dtype=torch.bfloat16 if args.bf16 else torch.float32)
    def sync():
        torch.cuda.synchronize()
        if world > 1: dist.barrier(); torch.cuda.synchronize()
    ret = tr.rollout(); tr.update()  # warm-up (cuDNN autotune, allocator)
    sync()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_roll = t_upd = 0.0
    rets = []
    for it in range(args.iters):
        base = (it + 1) * world * args.envs + rank * args.envs + 1
        e[0].record()
        rets.append(tr.rollout(episodes=torch.arange(args.envs) + base))
        e[1].record()
        stats = tr.update()
        e[2].record()
        sync()
        print("rank %d iter %d: rollout %.1f ms update %.1f ms" % (rank, it, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])), flush=True)
        t_roll += e[0].elapsed_time(e[1]); t_upd += e[1].elapsed_time(e[2])
    tt = torch.tensor([t_roll, t_upd], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_roll, t_upd = tt.tolist()
    if rank == 0:
        steps = world * args.envs * env.T * args.iters
        print(json.dumps({
            "metric": "train_env_steps_per_sec", "value": steps / ((t_roll + t_upd) * 1e-3), "unit": "env-steps/s",
            "rollout_env_steps_per_sec": steps / (t_roll * 1e-3), "n_gpus": world, "envs_per_gpu": args.envs,
            "agents": args.agents, "iters": args.iters, "data_passes": args.passes, "minibatch": args.minibatch,
            "ms_rollout_per_iter": t_roll / args.iters, "ms_update_per_iter": t_upd / args.iters,
            "compute_dtype": "bf16 autocast" if args.bf16 else "fp32", "mean_return": float(torch.stack(rets).mean()),
            "critic_loss": float(stats["critic_loss"]), "actor_loss": float(stats["actor_loss"]),
            "model_tflops": tr.flops_per_update() * args.iters / ((t_roll + t_upd) * 1e-3) / 1e12 * world,
            "grad_allreduce_bytes_per_step": 4 * (2275846 + 2307846) if world > 1 else 0}))
    if world > 1: dist.destroy_process_group()

if __name__ == "__main__":
    main()
