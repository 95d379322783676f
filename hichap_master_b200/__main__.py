"""``python -m hichap_master_b200 matrix ...`` -- the ``matrix`` sub-command of the reference CLI
(scripts/hichap:384-427 flags, :1049-1101 dispatch) on the B200 kernels.  Only this sub-command
exists: the other HiCHap stages are outside this package's scope."""
from __future__ import annotations

import argparse
import logging
import os
import sys


def getargs(argv=None):
    parser = argparse.ArgumentParser(prog="hichap_master_b200",
                                     description="B200 implementation of HiCHap's matrix stage")
    sub = parser.add_subparsers(dest="subcommand")
    m = sub.add_parser("matrix", help="Contact Matrix Construction",
                       formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    m.add_argument("-w", "--workspace", default=".", help="workspace directory (log file location)")
    m.add_argument("-log", "--logFile", default="hichap.log")
    m.add_argument("-b", "--bedPath", nargs="+", help="filtered bed path(s); several paths = replicates to merge")
    m.add_argument("-o", "--out", help="Output Folder.")
    m.add_argument("-N", "--NonAllelic", action="store_true", default=False,
                   help="if set, running Traditional HiC Matrix Pipeline")
    m.add_argument("-gs", "--genomeSize", help="genomeSize file Path.")
    m.add_argument("-wR", "--wholeRes", nargs="+", type=int, default=None,
                   help="Genome-Wide Matrix Resolution. default : None (only intra-chromosome matrices). Unit: bp")
    m.add_argument("-lR", "--localRes", nargs="+", type=int, default=[500000, 40000],
                   help="Intra-Chromosome Matrix Resolution. Unit : bp")
    m.add_argument("-ratio", "--ImputationRatio", type=float, default=0.9)
    m.add_argument("-min", "--ImputationMin", type=int, default=2)
    m.add_argument("-region", "--ImputationRegion", type=int, default=10000000)
    m.add_argument("-C", "--chroms", nargs="*", default=["#", "X"],
                   help='chromosome labels to include; "#" = numerical labels; no argument = all')
    argv = sys.argv[1:] if argv is None else argv
    if not argv or (argv[0] == "matrix" and len(argv) == 1):
        argv = list(argv) + ["-h"]
    return parser.parse_args(argv)


def run(argv=None):
    args = getargs(argv)
    if args.subcommand != "matrix":
        raise SystemExit("only the 'matrix' sub-command is implemented")
    logging.addLevelName(21, "main")                       # scripts/hichap:463-479
    logging.basicConfig(filename=os.path.join(args.workspace, args.logFile), level=21,
                        format="%(name)-25s %(levelname)-7s @ %(asctime)s: %(message)s")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:                       # launched with torchrun: one rank per GPU
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    from .matrixBuilding import HaplotypeMatrixConstruction, TraditionalMatrixConstruction
    whole = args.wholeRes or []
    os.makedirs(args.out, exist_ok=True)
    if args.NonAllelic:
        TraditionalMatrixConstruction(OutPath=args.out, RepPath=args.bedPath, genomeSize=args.genomeSize,
                                      wholeRes=whole, localRes=args.localRes, chroms=args.chroms)
    else:
        HaplotypeMatrixConstruction(OutPath=args.out, RepPath=args.bedPath, genomeSize=args.genomeSize,
                                    wholeRes=whole, localRes=args.localRes, Imputation_ratio=args.ImputationRatio,
                                    Imputation_min=args.ImputationMin, Imputation_region=args.ImputationRegion,
                                    chroms=args.chroms)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    run()
