"""Command line with the reference's flags (scripts/peakachu:5-89) for ``score_chromosome``,
``score_genome`` and ``depth``. ``train`` and ``pool`` are host-side tools of the reference and
are not part of this path; run the reference's own for those (``pool`` consumes our bedpe as
is; ``peakachu_b200.trainUtils.buildmatrix`` provides the training features).
"""
import argparse
import sys


def getargs(argv=None):
    parser = argparse.ArgumentParser(description="Peakachu loop scoring on B200 (sm_100a).",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    subparsers = parser.add_subparsers(dest="subcommands")
    subchrom = subparsers.add_parser("score_chromosome",
                                     help="Calculate interaction probability per pixel for a chromosome")
    subgen = subparsers.add_parser("score_genome",
                                   help="Calculate interaction probability per pixel for the whole genome")
    subdepth = subparsers.add_parser("depth", help="Calculate the total number of intra-chromosomal chromatin "
                                                   "contacts and select the most appropriate pre-trained model.")
    from . import calculate_depth, score_chromosome, score_genome
    subchrom.set_defaults(func=score_chromosome.main)
    subgen.set_defaults(func=score_genome.main)
    subdepth.set_defaults(func=calculate_depth.main)
    subdepth.add_argument("-p", "--path", help="Path to a .cool URI string")
    subdepth.add_argument("--min-dis", default=0, type=int,
                          help="Only count reads with genomic distance (in base pairs) greater than this value.")
    subdepth.add_argument("--device", type=int, default=None, help="CUDA device. Not a reference flag.")
    for i in (subchrom, subgen):
        i.add_argument("-r", "--resolution", help="Resolution in bp (default 10000)", type=int, default=10000)
        i.add_argument("-p", "--path", help="Path to a .cool URI string")
        i.add_argument("--clr-weight-name", default="weight",
                       help='The name of the weight column in your Cooler URI for normalizing the contact '
                            'signals. Specify it to "raw" if you want to use the raw signals.')
    subchrom.add_argument("-C", "--chrom", help="Chromosome label. Only contact data within the specified "
                                                "chromosome will be considered.")
    subgen.add_argument("-C", "--chroms", nargs="*", default=["#", "X"],
                        help='List of chromosome labels. "#" stands for chromosomes with numerical labels. '
                             '"--chroms" with zero argument will include all chromosome data.')
    for i in (subchrom, subgen):
        i.add_argument("-m", "--model", type=str, help="Path to pickled model file.")
        i.add_argument("-l", "--lower", type=int, default=6,
                       help="Lower bound of distance between loci in bins (default 6).")
        i.add_argument("-u", "--upper", type=int, default=300,
                       help="Upper bound of distance between loci in bins (default 300).")
        i.add_argument("--minimum-prob", type=float, default=0.5,
                       help="Only output pixels with probability score greater than this value (default 0.5)")
        i.add_argument("-O", "--output", help="Output file name.")
        i.add_argument("--device", type=int, default=None,
                       help="CUDA device (default: LOCAL_RANK, else 0). Not a reference flag.")
    commands = list(sys.argv[1:] if argv is None else argv)
    if (not commands) or (commands[0] in ("score_chromosome", "score_genome", "depth") and len(commands) == 1):
        commands.append("-h")
    return parser.parse_args(commands), commands


def run(argv=None):
    args, commands = getargs(argv)
    if commands[0] not in ("-h", "--help"):
        args.func(args)


if __name__ == "__main__":
    run()
