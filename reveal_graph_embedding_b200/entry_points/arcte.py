"""The `arcte` console script with the reference's flags (entry_points/arcte.py:12-84)."""
import argparse

import numpy as np
import scipy.sparse as spsp

from ..embedding.arcte.arcte import arcte
from ..io import read_adjacency_matrix, write_features


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-i", "--input", dest="input_edge_list_path", type=str, required=True,
                        help="This is the file path of the graph in edge list format.")
    parser.add_argument("-o", "--output", dest="output_feature_path", type=str, required=True,
                        help="This is the file path of the output features.")
    parser.add_argument("-s", "--separator", dest="separator", type=str, required=False, default="\t",
                        help="The character(s) separating the values in the edge list (default is tab).")
    parser.add_argument("-u", "--undirected", dest="undirected", type=bool, required=False, default=False,
                        help="Also create the reciprocal edge for each edge in edge list.")
    parser.add_argument("-r", "--rho", dest="restart_probability", type=float, required=False, default=0.1,
                        help="The restart probability for the vertex-centric PageRank calculation.")
    parser.add_argument("-e", "--epsilon", dest="epsilon_threshold", type=float, required=False,
                        default=1.0e-05, help="The tolerance for calculating vertex-centric PageRank values.")
    parser.add_argument("-nt", "--tasks", dest="number_of_tasks", type=int, required=False, default=None,
                        help="The number of GPUs to use (the reference: parallel tasks).")
    args = parser.parse_args(argv)

    adjacency_matrix, node_to_id = read_adjacency_matrix(file_path=args.input_edge_list_path,
                                                         separator=args.separator,
                                                         undirected=args.undirected)
    # entry_points/arcte.py:70-71: make sure the matrix is symmetric
    adjacency_matrix = spsp.csr_matrix(adjacency_matrix)
    adjacency_matrix = (adjacency_matrix + adjacency_matrix.transpose()) / 2

    features = arcte(adjacency_matrix=adjacency_matrix, rho=args.restart_probability,
                     epsilon=args.epsilon_threshold, number_of_threads=args.number_of_tasks)
    features = spsp.csr_matrix(features)
    write_features(file_path=args.output_feature_path, features=features, separator=args.separator,
                   node_to_id=node_to_id)


if __name__ == "__main__":
    main()
