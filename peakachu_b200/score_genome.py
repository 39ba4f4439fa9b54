"""``peakachu score_genome`` on the CUDA path (mirror of score_genome.py:3-84).

Chromosomes are independent (score_genome.py:46-84 shares only the model), so with
more than one GPU they are sharded greedily by band pixels (``shard.plan``); each
rank scores its shard and rank 0 gathers records on the host and writes them in the
reference's order (file order of chromosomes, then x, y). No device collective.
"""


def select_chromosomes(chromnames, chroms):
    """score_genome.py:39-44."""
    queue = []
    for key in chromnames:
        chromlabel = key.lstrip("chr")
        if (not chroms) or (chromlabel.isdigit() and "#" in chroms) or (chromlabel in chroms):
            queue.append(key)
    return queue


def main(args):
    import os

    from . import coolio, shard
    from .forest import load_model

    rank, world = shard.rank_world()
    if rank == 0 and os.path.exists(args.output):          # score_genome.py:11-12
        os.remove(args.output)

    flat, _ = load_model(args.model)                       # :14
    correct = False if args.clr_weight_name.lower() == "raw" else args.clr_weight_name   # :17-20
    Lib = coolio.open_map(args.path)                       # :28-31
    queue = select_chromosomes(Lib.chromnames[:], args.chroms)   # :39-44

    text = shard.score_chromosomes(Lib, queue, flat, correct=correct, lower=args.lower, upper=args.upper,
                                   res=args.resolution, min_prob=args.minimum_prob,
                                   device=getattr(args, "device", None), verbose=True)
    if rank == 0:
        with open(args.output, "a") as out:                # :83-84, chromosome by chromosome
            for key in queue:
                out.write(text[key])
