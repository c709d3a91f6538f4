"""Command-line drop-ins for the reference's pgen_esm.py / pgen_msa.py / pgen_msa_revised.py (same flags, same
TSV grammar, same output files), running the Gibbs loop on the B200 engine."""
import sys


def spec_args(text):
    """The dict literal of a specification line.  `/root/reference/src/pgen/pgen_esm.py:25` uses bare `eval`; here the
    text is parsed with `ast` and only literals are evaluated -- plus the two spellings of the documented `burnin`
    value, `float('inf')` and `inf` (and their negatives), and `range(...)` / `list(range(...))` of integer literals for
    `indexes`, which `ast.literal_eval` alone would reject.  Nothing in
    the line can call a function or reach an attribute."""
    import ast

    class _Inf(ast.NodeTransformer):
        def visit_Call(self, node):
            node.args = [self.visit(a) for a in node.args]
            name = node.func.id if isinstance(node.func, ast.Name) else None
            ints = [a.value for a in node.args if isinstance(a, ast.Constant) and type(a.value) is int]
            if name == "range" and not node.keywords and 1 <= len(node.args) <= 3 and len(ints) == len(node.args):
                elts = [ast.Constant(v) for v in range(*ints)]     # `indexes` is often written list(range(a, b))
                return ast.copy_location(ast.List(elts=elts, ctx=ast.Load()), node)
            if name == "list" and not node.keywords and len(node.args) == 1 and isinstance(node.args[0], (ast.List, ast.Tuple)):
                return ast.copy_location(ast.List(elts=node.args[0].elts, ctx=ast.Load()), node)
            if (isinstance(node.func, ast.Name) and node.func.id == "float" and len(node.args) == 1 and not node.keywords
                    and isinstance(node.args[0], ast.Constant) and isinstance(node.args[0].value, str)):
                return ast.copy_location(ast.Constant(float(node.args[0].value)), node)
            raise ValueError("function calls are not allowed in sampler arguments: " + text)

        def visit_Name(self, node):
            if node.id in ("inf", "Infinity"):
                return ast.copy_location(ast.Constant(float("inf")), node)
            if node.id == "nan":
                return ast.copy_location(ast.Constant(float("nan")), node)
            raise ValueError("names are not allowed in sampler arguments: " + node.id)

    tree = _Inf().visit(ast.parse(text.strip(), mode="eval"))
    value = ast.literal_eval(ast.fix_missing_locations(tree))
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
