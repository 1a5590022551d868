"""Planner-side node record.

Parity: reference ``ravnest/operations/node.py:1-35``.  Adds an optional
``gpu`` ordinal (B200 topology mode) and reads parameter names from the
stage artefact written by our splitter instead of a TorchScript module.
"""
import os

from ..utils.serialization import load_submodel


class Node:
    def __init__(self, node_id=None, address=None, benchmarks=None, gpu=None):
        self.node_id = node_id
        self.address = address
        self.benchmarks = benchmarks or {'ram': 0, 'bandwidth': 1}
        self.gpu = gpu
        self.ring_ids = {}
        self.address_to_param = {}
        self.split_quota = None
        self.cluster_id = None
        self.trainable_param_keys = []
        self.forward_target_host = None
        self.forward_target_port = None
        self.backward_target_host = None
        self.backward_target_port = None
        self.cluster_length = 1
        self.submodel = None
        self.stage_index = None

    @property
    def template_path(self):
        return 'node_data/cluster_{}/{}/'.format(self.cluster_id, self.address)

    def set_cluster_id(self, cid):
        self.cluster_id = cid

    def set_submodel(self, root='.'):
        self.submodel = load_submodel(os.path.join(root, self.template_path, 'submod.pt'))

    def set_trainable_parameter_keys(self):
        self.trainable_param_keys = [n for n, p in self.submodel.named_parameters() if p.requires_grad]

    def __repr__(self):
        return "Node({}, cluster={}, addr={}, rings={}, addr2param={})".format(
            self.node_id, self.cluster_id, self.address, self.ring_ids, self.address_to_param)
