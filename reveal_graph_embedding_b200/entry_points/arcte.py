"""The `arcte` console script: same flags as the reference's (entry_points/arcte.py:16-45),
edge list in, feature list out, with the text I/O and the extraction running in the native library."""
import argparse

import scipy.sparse as spsp

from ..embedding.arcte.arcte import arcte
from ..io import read_adjacency_matrix, write_features

# (short flag, long flag, destination, type, default, required, help): the reference's command line
_FLAGS = (
    ("-i", "--input", "input_edge_list_path", str, None, True, "edge list: one `source<sep>target<sep>weight` row per edge"),
    ("-o", "--output", "output_feature_path", str, None, True, "where the `node<sep>community<sep>value` rows go"),
    ("-s", "--separator", "separator", str, "\t", False, "field separator of both files (default: tab)"),
    ("-u", "--undirected", "undirected", "flag", False, False, "add the reverse of every listed edge (bare `-u`, or `-u true|false`)"),
    ("-r", "--rho", "restart_probability", float, 0.1, False, "restart probability of the absorbing walks (default 0.1)"),
    ("-e", "--epsilon", "epsilon_threshold", float, 1.0e-05, False, "push threshold epsilon (default 1e-5)"),
    ("-nt", "--tasks", "number_of_tasks", int, None, False, "GPUs to use (the reference: worker processes); default all"),
)


def _to_bool(text):
    v = str(text).strip().lower()
    if v in ("1", "true", "t", "yes", "y", "on"):
        return True
    if v in ("0", "false", "f", "no", "n", "off", ""):
        return False
    raise argparse.ArgumentTypeError("expected true or false, got %r" % (text,))


def main(argv=None):
    parser = argparse.ArgumentParser(description="ARCTE community features of a graph, on B200 GPUs.")
    for short, long_, dest, typ, default, required, text in _FLAGS:
        if typ == "flag":
            # the reference declares type=bool (entry_points/arcte.py:27-29), which takes a value and reads ANY
            # non-empty string, "False" included, as True; here both the bare flag and an explicit value work
            parser.add_argument(short, long_, dest=dest, nargs="?", const=True, default=default, type=_to_bool, help=text)
        else:
            parser.add_argument(short, long_, dest=dest, type=typ, default=default, required=required, help=text)
    args = parser.parse_args(argv)

    graph, node_to_id = read_adjacency_matrix(file_path=args.input_edge_list_path, separator=args.separator,
                                              undirected=args.undirected)
    graph = spsp.csr_matrix(graph)
    graph = (graph + graph.transpose()) / 2        # entry_points/arcte.py:70-71: symmetrise
    features = arcte(graph, args.restart_probability, args.epsilon_threshold, args.number_of_tasks)
    write_features(file_path=args.output_feature_path, features=spsp.csr_matrix(features),
                   separator=args.separator, node_to_id=node_to_id)


if __name__ == "__main__":
    main()
