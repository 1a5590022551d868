"""One pipeline replica as the planner sees it.

A cluster is an ordered chain of planner nodes, each owning one stage of a complete copy of the model; several clusters train the
same model on different data and average parameters ring-wise.  Field and method names follow the reference's planner record
(``ravnest/operations/cluster.py``) because ``operations/utils.py`` and user scripts address them by name; the record itself is a
dataclass here and also knows how to wire its chain and to describe itself for logs.
"""
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional


@dataclass
class Cluster:
    cid: int
    nodes: "OrderedDict[int, object]" = field(default_factory=OrderedDict)     # node_id -> planner Node, in pipeline order
    rams: List[float] = field(default_factory=list)                               # per-member RAM, same order (split quotas)
    total_ram: float = 0
    total_speed: float = 0
    splits: Optional[List[float]] = None                                          # share of the model per member
    all_param_to_ring: Dict[str, int] = field(default_factory=dict)

    # ------------------------------------------------------------------ membership
    @property
    def size(self) -> int:
        return len(self.nodes)

    def add_node(self, node) -> None:
        """Append ``node`` as the next stage of this replica."""
        node.set_cluster_id(self.cid)
        self.nodes[node.node_id] = node
        self.rams.append(node.benchmarks['ram'])

    def ordered_nodes(self) -> list:
        return list(self.nodes.values())

    def wire(self) -> None:
        """Tell every member who precedes and who follows it (forward / backward targets, stage index)."""
        chain = self.ordered_nodes()
        for position, node in enumerate(chain):
            node.link(previous=chain[position - 1] if position > 0 else None,
                      following=chain[position + 1] if position + 1 < len(chain) else None, stage_index=position)

    # ------------------------------------------------------------------ quotas and rings
    def assign_split_quotas_to_nodes(self) -> None:
        if self.splits is None:
            raise ValueError('cluster {} has no split quotas yet'.format(self.cid))
        for node, quota in zip(self.nodes.values(), self.splits):
            node.split_quota = quota

    def set_ringwise_params(self) -> None:
        """For the cluster that defines the rings (the largest one): trainable parameter name -> id of the ring its stage owns."""
        self.all_param_to_ring = {key: next(iter(node.ring_ids)) for node in self.nodes.values() for key in node.trainable_param_keys}

    def describe(self) -> Dict[str, object]:
        return {'cluster': self.cid, 'members': [n.address for n in self.nodes.values()], 'splits': self.splits,
                'total_ram': self.total_ram}

    def __repr__(self) -> str:
        return 'Cluster({}, nodes={}, splits={})'.format(self.cid, list(self.nodes), self.splits)
