"""Planner-side record of one provider.

What the planner knows about a host / GPU before any runtime exists: its address and benchmarks, the cluster and stage it was
assigned to, its pipeline neighbours, the rings it takes part in and which parameter ranges go to which ring peer.  Attribute
names follow the reference (``ravnest/operations/node.py``) since they are serialised into ``node_data/nodes/node_<id>.json``;
this record additionally derives its role from its wiring, carries a GPU ordinal for grid topologies and reads parameter names
from the stage artefact written by our splitter (a pickled GraphModule, not TorchScript).
"""
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

from ..utils.serialization import load_submodel


@dataclass
class Node:
    node_id: Optional[int] = None
    address: Optional[str] = None                          # "host:port"
    benchmarks: Dict[str, float] = None                    # {'ram': GB, 'bandwidth': Mbps}
    gpu: Optional[int] = None                              # device ordinal on a GPU grid
    cluster_id: Optional[int] = None
    stage_index: Optional[int] = None
    cluster_length: int = 1
    split_quota: Optional[float] = None
    ring_ids: Dict[int, Optional[str]] = field(default_factory=dict)           # ring id -> first parameter of the ring
    address_to_param: Dict[str, str] = field(default_factory=dict)             # ring-peer address -> last parameter sent there
    trainable_param_keys: List[str] = field(default_factory=list)
    forward_target_host: Optional[str] = None
    forward_target_port: Optional[str] = None
    backward_target_host: Optional[str] = None
    backward_target_port: Optional[str] = None
    submodel: object = None

    def __post_init__(self):
        if self.benchmarks is None:
            self.benchmarks = {'ram': 0, 'bandwidth': 1}

    # ------------------------------------------------------------------ placement
    def set_cluster_id(self, cid: int) -> None:
        self.cluster_id = cid

    @property
    def template_path(self) -> str:
        """Folder (relative to the job root) that holds this node's stage artefact and routing templates."""
        return 'node_data/cluster_{}/{}/'.format(self.cluster_id, self.address)

    def link(self, previous: Optional['Node'], following: Optional['Node'], stage_index: int) -> None:
        self.stage_index = stage_index
        if following is not None:
            self.forward_target_host, self.forward_target_port = following.address.split(':')
        if previous is not None:
            self.backward_target_host, self.backward_target_port = previous.address.split(':')

    @property
    def role(self) -> str:
        """root / stem / leaf / solo, from which neighbours exist."""
        sends_on, sends_back = self.forward_target_host is not None, self.backward_target_host is not None
        if sends_on and sends_back:
            return 'stem'
        if sends_on:
            return 'root'
        return 'leaf' if sends_back else 'solo'

    # ------------------------------------------------------------------ stage artefact
    def set_submodel(self, root: str = '.') -> None:
        self.submodel = load_submodel(os.path.join(root, self.template_path, 'submod.pt'))

    def set_trainable_parameter_keys(self) -> None:
        self.trainable_param_keys = [name for name, p in self.submodel.named_parameters() if p.requires_grad]

    def __repr__(self) -> str:
        return 'Node({}, cluster={}, stage={}, addr={}, rings={})'.format(self.node_id, self.cluster_id, self.stage_index, self.address,
                                                                      sorted(self.ring_ids))
