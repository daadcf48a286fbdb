from .gat import GATConvB200, GATEncoderB200, gat_edge_softmax_aggregate
