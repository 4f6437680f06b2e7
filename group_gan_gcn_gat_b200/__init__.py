"""B200-native sgan social-interaction ops behind the reference's module API.

Drop-in for the hot path of peaceminusones/Group-GAN-GCN-GAT (sgan/models.py): PoolHiddenNet,
GraphAttentionLayer/GAT/GATEncoder, GCN/GCNModule and the TrajectoryGenerator/Discriminator wiring,
computed by hand-written sm_100a kernels in lib/libsgx_b200.so (C ABI: include/sgx.h).
"""
__version__ = '0.1.0'
