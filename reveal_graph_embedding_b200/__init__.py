"""B200-native ARCTE feature extractor.

Drop-in for the ARCTE path of MKLab-ITI/reveal-graph-embedding:

    from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
    X = arcte(adjacency_matrix, rho, epsilon, number_of_threads)

Everything below that call runs in hand-written CUDA kernels for sm_100a
(libarcte_cuda.so, C ABI in include/arcte_cuda.h).  There is no CPU fallback:
importing works anywhere, but any compute call without the built library and a
B200 raises.
"""
__version__ = "0.1.0"
