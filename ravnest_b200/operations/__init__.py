"""Offline planner (reference ``ravnest/operations/``)."""
from .cluster import Cluster
from .genetic import genetic_algorithm, calculate_fitness
from .node import Node as PlannerNode
from .splitter import (split_model, trace_model, choose_split_points, choose_split_points_by_cost, estimate_costs, measure_costs,
                       split_graph, StageTracer, SplitResult)
from .utils import (clusterize, spawn_node_pool, cluster_formation, grid_formation, round_percentages,
                    calculate_cluster_power, calculate_split_percentages, split_model_equal, get_memory_reqs,
                    delete_all_folders, equal_proportions, default_gpu_configs, split_model_on_proportions, remake_proportions,
                    split_on_proportions, CustomTracer, rebalanced_proportions)

__all__ = ['Cluster', 'genetic_algorithm', 'calculate_fitness', 'PlannerNode', 'split_model', 'trace_model',
           'choose_split_points', 'split_graph', 'StageTracer', 'SplitResult', 'clusterize', 'spawn_node_pool',
           'cluster_formation', 'grid_formation', 'round_percentages', 'calculate_cluster_power',
           'calculate_split_percentages', 'split_model_equal', 'get_memory_reqs', 'delete_all_folders',
           'equal_proportions', 'default_gpu_configs', 'split_model_on_proportions', 'remake_proportions', 'split_on_proportions',
           'CustomTracer', 'rebalanced_proportions', 'choose_split_points_by_cost', 'estimate_costs', 'measure_costs']
