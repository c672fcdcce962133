"""Stand-in package for `torch_geometric` 2.2.0 (only `nn.knn_graph` / `nn.radius_graph`). Test infrastructure."""
