"""``peakachu score_chromosome`` on the CUDA path (mirror of score_chromosome.py:3-71).

Same ``args`` namespace as the reference: path, chrom, model, lower, upper,
minimum_prob, output, resolution, clr_weight_name (+ optional ``device``).
"""


def main(args):
    import os

    from . import coolio
    from .forest import load_model
    from .scoreUtils import Chromosome, DeviceForest

    if os.path.exists(args.output):                       # score_chromosome.py:11-12
        os.remove(args.output)

    # :14 -- unpickling a 100-tree forest is ~0.1 s of interpreter time; the file's pixel columns are inflated by the
    # library's host threads meanwhile (H5Cool keeps the fetched chromosome for from_map below)
    import threading
    loaded = {}

    def _load():
        try:
            loaded["flat"] = load_model(args.model)[0]
        except BaseException as e:                        # re-raised on the calling thread
            loaded["error"] = e
    loader = threading.Thread(target=_load, name="pk-model-load")
    loader.start()
    map_error = None
    try:
        Lib = coolio.open_map(args.path)                  # :33-34 (.hic is outside this path)
        if hasattr(Lib, "prefetch"):
            Lib.prefetch(args.chrom)
    except Exception as e:
        map_error = e
    loader.join()
    if "error" in loaded:                                 # the reference fails on the model first
        raise loaded["error"]
    if map_error is not None:
        raise map_error
    flat = loaded["flat"]
    correct = False if args.clr_weight_name.lower() == "raw" else args.clr_weight_name   # :17-20
    width = flat.width                                    # :23
    device = int(getattr(args, "device", 0) or 0)

    ccname = args.chrom
    cikada = "chr" + ccname.lstrip("chr")                 # :37-38

    from .shard import map_weights
    weights, pweights = map_weights(Lib, ccname, correct)         # :44 (a divisive_weights column is inverted for the values)
    forest = DeviceForest.of(flat, device)
    # replaces :42-43 (matrix fetch + tocsr): the chromosome's pixel columns as the reader stores them
    X = Chromosome.from_map(Lib, ccname, weights, forest, lower=args.lower, upper=args.upper, cname=cikada,
                            res=args.resolution, width=width, device=device)
    if pweights is not None:
        X.set_poisson_weights(pweights)
    result, R = X.score(thre=args.minimum_prob)           # :70
    X.writeBed(args.output, result, R)                    # :71
    X.close()
