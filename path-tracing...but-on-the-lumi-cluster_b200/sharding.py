"""Frame sharding over GPUs (reference main.cc:78-102 made parallel: a frame is a pure function of
its index, scene.cc:271, so whole frames are the unit and ranks never exchange render data).

Two policies, both used with one process (or host thread) per GPU:
  * strided:  rank r of N takes frames r, r+N, r+2N, ... — static, no communication at all; used
              by bench.py under torchrun and by the gloo tests.
  * dynamic:  a shared counter hands out the next frame (host/ptgpu_main.cc uses an atomic in one
              process); per-frame cost varies ~7x over the animation, so this balances better.
"""


def strided_frames(n_frames, rank, world, begin=0, step=1):
    """Frames of `rank`: begin + step*(rank + k*world) for k = 0, 1, ... below n_frames."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world: %r/%r" % (rank, world))
    if step < 1:
        raise ValueError("step must be >= 1")
    return list(range(begin + step * rank, n_frames, step * world))


def bench_frame(step, rank, steps, available):
    """Snapshot frame rendered by `rank` at benchmark step `step` of `steps` (bench.py). The run's frame
    list is available[j % len] for j < steps; rank r walks it rotated by r. Every rank therefore renders the
    same MULTISET of frames whatever the rank count, so throughput ratios between rank counts measure
    scaling and not the frame mix (per-frame cost varies ~5x over the animation), while at any moment
    the ranks render different frames."""
    if not available:
        raise ValueError("no snapshot frames")
    if steps < 1:
        raise ValueError("steps must be >= 1")
    return available[((step + rank) % steps) % len(available)]


def check_partition(assignments, n_frames, begin=0, step=1):
    """True iff the per-rank frame lists cover every frame exactly once."""
    want = list(range(begin, n_frames, step))
    got = sorted(f for a in assignments for f in a)
    return got == want


def animation_seconds(per_frame_seconds, n_frames, world):
    """Wall-clock estimate of a whole animation from a mean per-frame time measured on one rank."""
    return per_frame_seconds * n_frames / float(world)
