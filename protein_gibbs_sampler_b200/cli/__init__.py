"""Command-line drop-ins for the reference's pgen_esm.py / pgen_msa.py / pgen_msa_revised.py (same flags, same
TSV grammar, same output files), running the Gibbs loop on the B200 engine."""
import sys


def spec_args(text):
    """The dict literal of a specification line (`/root/reference/src/pgen/pgen_esm.py:25` uses bare `eval`).
    Evaluated without builtins; `float('inf')` / `inf` -- the documented `burnin` values -- stay available."""
    value = eval(text, {"__builtins__": {}}, {"float": float, "int": int, "inf": float("inf"), "range": range,
                                               "list": list})
    if not isinstance(value, dict):
        raise ValueError("sampler arguments must be a dict literal, got: " + text)
    return value


def spec_lines(input_h, echo_h, n_fields, what, complain=True):
    """The runs of a specification file: non-empty lines split at tabs.  A line with the wrong number of fields is
    reported (like the reference scripts do) and skipped; accepted lines are echoed to stdout and to `echo_h`."""
    for raw in input_h:
        fields = raw.strip().split("\t")
        if fields == [""]:
            continue
        if len(fields) != n_fields:
            if complain:
                print(f"Expected {n_fields} values in specification file ({what}), got {len(fields)}")
                print("\t".join(fields))
            continue
        print("\t".join(fields))
        print("\t".join(fields), file=echo_h)
        yield fields


def add_weight_flags(parser):
    parser.add_argument("--checkpoint", default=None,
                        help="path to a fair-esm checkpoint (.pt) for --model.  Without it the model runs on seeded "
                             "synthetic weights (no pretrained weights can be downloaded in this environment).")
    parser.add_argument("--weights_seed", type=int, default=0, help="seed of the synthetic weights.")


def build_model(model_map, args):
    cls = model_map[args.model]
    if args.checkpoint is None:
        print("warning: no --checkpoint given, using synthetic weights (seed %d)" % args.weights_seed, file=sys.stderr)
        return cls(seed=args.weights_seed)
    return cls(checkpoint=args.checkpoint)
