"""Multi-GPU parity check, launched with torchrun (one process per GPU, NCCL):

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_check.py

(1) experiments sharded over ranks reproduce the unsharded run; (2) a tempering ladder sharded
over ranks (all-gather of energies per swap step) equals the single-rank ladder; (3) a lattice
in row strips with NCCL halo exchange equals one strip.  Rank 0 prints MULTI_GPU_OK."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import pyisingmontecarlo_b200 as pkg
    from pyisingmontecarlo_b200 import _native as nat

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = nat.Context.get(local)

    # (1) sharded experiments: 64 per rank, offsets keep the Philox streams global
    g = nat.Graph.torus(ctx, (8, 8, 8), j0=1.0, pmj=True, j_seed=4)
    betas = np.linspace(0.2, 1.0, 6)
    mine = nat.Sim(g, 64, 99, replica_offset=64 * rank)
    mine.sweeps(betas)
    full = nat.Sim(g, 64 * world, 99)
    full.sweeps(betas)
    assert (full.states()[64 * rank:64 * (rank + 1)] == mine.states()).all()

    # (1b) the Lattice API with distributed = True: every rank returns the full, identical arrays
    sq = [((x * 8 + y, ((x + 1) % 8) * 8 + y), -1.0) for x in range(8) for y in range(8)] + \
         [((x * 8 + y, x * 8 + (y + 1) % 8), -1.0) for x in range(8) for y in range(8)]
    lat = pkg.Lattice(sq, seed_gen=5, device=local)
    e_full, s_full = lat.run_monte_carlo_annealing_and_get_energies([(0, 0.2), (6, 0.8)], 6, 100)
    lat.distributed = True
    e_sh, s_sh = lat.run_monte_carlo_annealing_and_get_energies([(0, 0.2), (6, 0.8)], 6, 100)
    assert e_sh.shape == (100, 6) and (s_sh == s_full).all() and (e_sh == e_full).all()
    en_s, st_s = lat.run_monte_carlo_sampling(0.4, 6, 70, None, 1, 2)
    lat.distributed = False
    en_f, st_f = lat.run_monte_carlo_sampling(0.4, 6, 70, None, 1, 2)
    assert (st_s == st_f).all() and (en_s == en_f).all()

    # (2) sharded tempering == single-rank tempering
    edges = [((x * 6 + y, ((x + 1) % 6) * 6 + y), -1.0) for x in range(6) for y in range(6)] + \
            [((x * 6 + y, x * 6 + (y + 1) % 6), -1.0) for x in range(6) for y in range(6)]
    ladder = np.linspace(0.2, 0.7, 2 * world + 1)
    lt = pkg.LatticeTempering(edges, seed=11, device=local)
    for b in ladder:
        lt.add_graph(0.0, 0.0, b)
    states, energies = lt.qmc_timesteps_sample(40, replica_swap_freq=3, sampling_freq=8)
    single = nat.Tempering(pkg.Lattice(edges, device=local).graph(), ladder, 11)
    st1, en1 = single.timesteps_sample(40, 3, 8)
    assert (states == st1).all() and np.array_equal(energies, en1), "sharded tempering differs"
    assert lt.get_total_swaps() == single.total_swaps() > 0
    # a sharded ladder checkpoints per rank and continues bit for bit
    import tempfile
    ckdir = [tempfile.mkdtemp() if rank == 0 else None]
    dist.broadcast_object_list(ckdir, src=0)
    ck = os.path.join(ckdir[0], "ladder.npz")
    lt.save_to_file(ck)
    dist.barrier()
    st_a, en_a = lt.qmc_timesteps_sample(24, replica_swap_freq=3, sampling_freq=8)
    lt2 = pkg.LatticeTempering.read_from_file(ck, device=local)
    st_b, en_b = lt2.qmc_timesteps_sample(24, replica_swap_freq=3, sampling_freq=8)
    assert (st_a == st_b).all() and np.array_equal(en_a, en_b), "restored sharded ladder differs"
    assert lt2.get_total_swaps() == lt.get_total_swaps()

    # (3) strips + NCCL halo exchange == one strip
    sweep_betas = [0.4, 0.5, 0.3, 0.44, 0.6]
    lats = [pkg.SingleLattice2D(256, 16 * world, seed=5, device=local, exchange_every=k) for k in (0, 2, 8)]
    for lat in lats:      # per-phase overlapped exchange, batches of 2 (partial last), one deep batch
        lat.sweeps(sweep_betas)
    for other in lats[1:]:
        assert (other.local_rows() == lats[0].local_rows()).all(), "exchange modes differ"
    lat = lats[2]
    e = lat.energy()
    whole = nat.Strip(ctx, 256, 16 * world, 0, 16 * world, -1.0, 5)
    for b in sweep_betas:
        for colour in (0, 1):
            whole.wrap_local(1 - colour)
            whole.phase(colour, b)
    assert (whole.rows()[lat.row_lo:lat.row_hi] == lat.local_rows()).all(), "strip decomposition differs"
    whole.wrap_local(1)
    nsat, _ = whole.observables()
    assert e == 2 * 256 * 16 * world - 2 * nsat
    dist.barrier()
    if rank == 0:
        print(f"MULTI_GPU_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
