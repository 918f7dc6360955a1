"""Packaging of the B200-native ARCTE extractor: registers the `arcte` console script like the reference's
setup.py:108-110.  The CUDA library is built in-tree first (`python -c "import __graft_entry__ as g; g.build()"`
or `make -C reveal_graph_embedding_b200/csrc`) and shipped as package data."""
from setuptools import find_packages, setup

setup(
    name="reveal-graph-embedding-b200",
    version="0.2.0",
    description="ARCTE community-feature extraction on B200 GPUs behind the reveal-graph-embedding API",
    packages=find_packages(include=["reveal_graph_embedding_b200", "reveal_graph_embedding_b200.*"]),
    package_data={"reveal_graph_embedding_b200": ["libarcte_cuda.so"]},
    python_requires=">=3.9",
    install_requires=["numpy", "scipy"],
    entry_points={"console_scripts": ["arcte=reveal_graph_embedding_b200.entry_points.arcte:main"]},
)
