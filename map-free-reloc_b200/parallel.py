"""Multi-GPU driver logic: the pair stream shards over ranks in contiguous blocks (pair order kept
inside a rank), weights are replicated, and the only collective is the final gather of fixed-size
pose records. Works with any torch.distributed backend (NCCL on the GPUs, gloo in
the CPU tests)."""
import torch
import torch.distributed as dist

RECORD_WIDTH = 14  # pair_id, inliers, R (9), t (3)


def shard_range(n_pairs, rank, world):
    """Contiguous block [start, stop) of the pair stream owned by ``rank`` (sizes differ by <= 1)."""
    base, rem = divmod(n_pairs, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_records(pair_ids, R, t, inliers):
    """[n] ids, [n,3,3], [n,3], [n] -> float64 [n, 14] records (float64 keeps int ids exact)."""
    n = len(pair_ids)
    rec = torch.zeros(n, RECORD_WIDTH, dtype=torch.float64, device=R.device)
    rec[:, 0] = torch.as_tensor(pair_ids, dtype=torch.float64, device=R.device)
    rec[:, 1] = inliers.to(torch.float64)
    rec[:, 2:11] = R.reshape(n, 9).to(torch.float64)
    rec[:, 11:14] = t.reshape(n, 3).to(torch.float64)
    return rec


def gather_records(rec, n_pairs_total):
    """All ranks contribute their [n_r, 14] records; returns the [n_pairs_total, 14] table ordered by
    pair id on every rank (ragged shards are padded to the largest one for the collective)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rec[torch.argsort(rec[:, 0])]
    world = dist.get_world_size()
    cap = (n_pairs_total + world - 1) // world
    padded = torch.full((cap, RECORD_WIDTH), float("nan"), dtype=rec.dtype, device=rec.device)
    padded[:, 0] = -1
    padded[: rec.shape[0]] = rec
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    table = torch.cat(out, 0)
    table = table[table[:, 0] >= 0]
    return table[torch.argsort(table[:, 0])]


def median_pose_errors(table, R_ref, t_ref):
    """Median rotation angle [rad] and translation distance [m] against reference poses, NaN rows
    (failed pairs, submission.py:48-49) excluded. Formulas: lib/utils/metrics.py:24-28."""
    R = table[:, 2:11].reshape(-1, 3, 3)
    t = table[:, 11:14]
    ok = ~(torch.isnan(R).any(-1).any(-1) | torch.isnan(t).any(-1))
    if ok.sum() == 0:
        return float("nan"), float("nan"), 0
    Rr = torch.as_tensor(R_ref, dtype=torch.float64)[ok]
    tr = torch.as_tensor(t_ref, dtype=torch.float64)[ok]
    c = ((torch.einsum("nij,nij->n", R[ok], Rr)) - 1.0) / 2.0
    ang = torch.acos(c.clamp(-1.0, 1.0))
    dt = (t[ok] - tr).norm(dim=1)
    return float(ang.median()), float(dt.median()), int(ok.sum())


def run_stream(n_pairs, batch, submit, drain, rank=0, world=1, device="cpu"):
    """Drives one rank's contiguous block of a stream of ``n_pairs`` pairs through a one-deep pipeline and returns the
    complete, pair-ordered record table [n_pairs, 14] on every rank (the multi-GPU replacement of the reference's
    sequential loop, submission.py:36-56 — pairs are independent, so the only collective is the final gather).

    ``submit(ids)`` enqueues one engine batch of exactly ``batch`` pair ids and returns the finished result of the batch
    submitted one call earlier (``None`` on the first call), ``drain()`` returns the last one; a result is
    ``(R [batch,3,3], t [batch,3], inliers [batch])`` (``mfr_b200.pipeline.RelocPipeline.submit_host`` / ``drain``).
    A ragged last batch is padded by repeating its final pair id; the padding rows are dropped again."""
    a, b = shard_range(n_pairs, rank, world)
    ids = list(range(a, b))
    chunks = [ids[k:k + batch] for k in range(0, len(ids), batch)]
    done, pending = [], []
    for chunk in chunks:
        prev = submit(chunk + [chunk[-1]] * (batch - len(chunk)))
        pending.append(chunk)
        if prev is not None:
            done.append((pending.pop(0), prev))
    if pending:
        done.append((pending.pop(0), drain()))
    assert not pending and sum(len(c) for c, _ in done) == len(ids)
    if done:
        R = torch.cat([torch.as_tensor(r[0])[: len(c)] for c, r in done]).to(device)
        t = torch.cat([torch.as_tensor(r[1])[: len(c)] for c, r in done]).to(device)
        n = torch.cat([torch.as_tensor(r[2])[: len(c)] for c, r in done]).to(device)
    else:
        R, t, n = torch.zeros(0, 3, 3, device=device), torch.zeros(0, 3, device=device), torch.zeros(0, dtype=torch.int32, device=device)
    return gather_records(pack_records(ids, R, t, n), n_pairs)
