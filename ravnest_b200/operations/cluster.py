"""Planner-side description of one pipeline replica ("cluster").

Parity: reference ``ravnest/operations/cluster.py:1-34``.  On a B200 box a
cluster is one pipeline replica: an ordered chain of GPUs, each owning one stage.
"""
from collections import OrderedDict


class Cluster:
    def __init__(self, cid):
        self.cid = cid
        self.nodes = OrderedDict()      # node_id -> PlannerNode, in pipeline order
        self.size = 0
        self.rams = []
        self.total_ram = 0
        self.total_speed = 0
        self.splits = None
        self.all_param_to_ring = {}

    def add_node(self, node):
        node.set_cluster_id(self.cid)
        self.nodes[node.node_id] = node
        self.rams.append(node.benchmarks['ram'])
        self.size += 1

    def ordered_nodes(self):
        return list(self.nodes.values())

    def set_ringwise_params(self):
        """Every trainable parameter of this (largest) cluster -> the ring id owned by its stage."""
        mapping = {}
        for node in self.nodes.values():
            ring_id = next(iter(node.ring_ids))
            for key in node.trainable_param_keys:
                mapping[key] = ring_id
        self.all_param_to_ring = mapping

    def assign_split_quotas_to_nodes(self):
        for quota, node in zip(self.splits, self.nodes.values()):
            node.split_quota = quota

    def __repr__(self):
        return "Cluster({}, nodes={}, splits={})".format(self.cid, list(self.nodes.keys()), self.splits)
