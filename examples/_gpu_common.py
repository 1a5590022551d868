"""Shared launcher for the GPU examples: one process per GPU under torchrun, rank == node id.

  python cluster_formation.py resnet50 4x2
  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/resnet50/provider.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ravnest_b200.runtime.dist_transport import init_distributed  # noqa: E402


def launch(make_node, make_trainer):
    rank, world, dev = init_distributed()
    node = make_node('node_{}'.format(rank), dev)
    trainer = make_trainer(node)
    trainer.train()
    trainer.finish(grace=0.2, timeout=60)           # root: drain and stop the pipeline; others: wait for the SHUTDOWN
    node.compute_session.flush_losses()
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
