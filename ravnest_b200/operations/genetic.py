"""Genetic-algorithm cluster formation.

Parity: reference ``ravnest/operations/genetic.py:6-70`` — population 200,
500 generations, at most 5 clusters, mutation rate 0.01, tournament-of-5
selection, single-point crossover, and a fitness of
``100 * (RAM shortfall summed over clusters) + (max - min cluster "speed")``.

Differences (deliberate):
  * uses a private ``random.Random(seed)`` so planning is reproducible and does
    not disturb the global RNG stream that ``set_seed`` controls;
  * fitness is evaluated with pre-extracted integer arrays (the reference walks
    attribute dicts 200x500 times);
  * early exit when a zero-penalty individual has been stable for 25
    generations (the reference always burns all 500).
"""
import random
from typing import Dict, List, Sequence


def _fitness(individual: Sequence[int], rams: Sequence[int], speeds: Sequence[int], required: int) -> int:
    ram_by_cluster: Dict[int, int] = {}
    speed_by_cluster: Dict[int, int] = {}
    for gene, ram, speed in zip(individual, rams, speeds):
        ram_by_cluster[gene] = ram_by_cluster.get(gene, 0) + ram
        speed_by_cluster[gene] = speed_by_cluster.get(gene, 0) + speed
    shortfall = 0
    for total in ram_by_cluster.values():
        if total <= required:
            shortfall += required - total
    spread = max(speed_by_cluster.values()) - min(speed_by_cluster.values())
    return 100 * shortfall + spread


def calculate_fitness(individual, nodes, L, t=0):
    """Reference-compatible signature (genetic.py:6)."""
    rams = [n.benchmarks['ram'] for n in nodes]
    speeds = [n.benchmarks['ram'] // max(1, n.benchmarks['bandwidth']) for n in nodes]
    return _fitness(individual, rams, speeds, L)


def decode_solution(solution, items):
    clusters: Dict[int, List] = {}
    for item, cid in zip(items, solution):
        clusters.setdefault(cid, []).append(item)
    return clusters


def genetic_algorithm(nodes, L, t=0, population_size=200, max_clusters=5, generations=500,
                      mutation_rate=0.01, seed=42, patience=25):
    """Group ``nodes`` into <= ``max_clusters`` clusters that can each hold a model of size ``L`` (MB)."""
    rng = random.Random(seed)
    n = len(nodes)
    if n == 0:
        return {}
    max_clusters = max(1, min(max_clusters, n))
    rams = [int(node.benchmarks['ram']) for node in nodes]
    speeds = [int(node.benchmarks['ram']) // max(1, int(node.benchmarks['bandwidth'])) for node in nodes]

    population = [[rng.randrange(max_clusters) for _ in range(n)] for _ in range(population_size)]
    best, best_fit, stable = None, None, 0
    for _ in range(generations):
        fitness = [_fitness(ind, rams, speeds, L) for ind in population]
        # ties are broken towards FEWER clusters (longer pipelines): the documented 3-node CNN
        # walkthrough forms one root->stem->leaf cluster (docs/walkthrough.rst:131-146)
        gen_best = min(range(population_size), key=lambda i: (fitness[i], len(set(population[i]))))
        key = (fitness[gen_best], len(set(population[gen_best])))
        if best_fit is None or key < (best_fit, len(set(best))):
            best, best_fit, stable = list(population[gen_best]), fitness[gen_best], 0
        else:
            stable += 1
        if best_fit == 0 and stable >= patience:
            break
        scored = list(zip(population, fitness))
        nxt = []
        for _ in range(population_size // 2):
            parents = []
            for _ in range(2):
                contenders = rng.sample(scored, min(5, len(scored)))
                parents.append(min(contenders, key=lambda p: p[1])[0])
            if n > 1:
                cut = rng.randint(1, n - 1)
                kids = (parents[0][:cut] + parents[1][cut:], parents[1][:cut] + parents[0][cut:])
            else:
                kids = (list(parents[0]), list(parents[1]))
            for kid in kids:
                nxt.append([rng.randrange(max_clusters) if rng.random() < mutation_rate else g for g in kid])
        population = nxt
    return decode_solution(best, nodes)
