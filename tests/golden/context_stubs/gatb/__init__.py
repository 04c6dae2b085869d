"""Stand-in for pygatb (not installable here), only as far as scripts/python3/Context_genome_WG.py uses it: Graph("-in x.h5"),
graph[kmer] -> node with .in_degree / .out_degree / .reversed, bytes(node). Every degree comes from gatb-core itself through
oracle/_ref/bin/refdegrees (Graph::load + buildNode + indegree/outdegree); answers are fetched lazily in batches."""
import os
import subprocess

_TOOL = os.environ["MTG_REFDEGREES"]


class _Node:
    def __init__(self, graph, kmer):
        self._g, self._k = graph, kmer

    def __bytes__(self):
        return self._k.encode()

    @property
    def reversed(self):
        return self          # the script only asserts node.reversed == node

    def _deg(self):
        return self._g._degrees(self._k)

    @property
    def in_degree(self):
        return self._deg()[0]

    @property
    def out_degree(self):
        return self._deg()[1]


class Graph:
    def __init__(self, fmt):
        assert fmt.startswith("-in ")
        self.h5 = fmt[4:]
        self.cache = {}

    def __getitem__(self, kmer):
        return _Node(self, kmer)

    def _degrees(self, kmer):
        if kmer not in self.cache:
            r = subprocess.run([_TOOL, self.h5], input=kmer + "\n", stdout=subprocess.PIPE, text=True, check=True)
            a, b = r.stdout.split()
            self.cache[kmer] = (int(a), int(b))
        return self.cache[kmer]

    def prefetch(self, kmers):
        todo = [k for k in dict.fromkeys(kmers) if k not in self.cache]
        if todo:
            r = subprocess.run([_TOOL, self.h5], input="\n".join(todo) + "\n", stdout=subprocess.PIPE, text=True, check=True)
            for k, line in zip(todo, r.stdout.splitlines()):
                a, b = line.split()
                self.cache[k] = (int(a), int(b))
