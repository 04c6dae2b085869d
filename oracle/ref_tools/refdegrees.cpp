// oracle/_ref helper (test infrastructure): the degrees pygatb's `graph[kmer].in_degree / out_degree` would return, from gatb-core
// itself: Graph::load on a reference .h5, buildNode(kmer string) = the node in the strand of the string, Graph::indegree /
// Graph::outdegree (debruijn/impl/Graph.hpp). Used to run the UNMODIFIED scripts/python3/Context_genome_WG.py without pygatb
// (tests/golden/make_context_fixture.py).    refdegrees x.h5 < kmers.txt   ->  one "in out" line per k-mer
#include <gatb/gatb_core.hpp>
#include <iostream>
#include <string>
int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: refdegrees graph.h5 < kmers\n"); return 2; }
    try {
        Graph graph = Graph::load(argv[1]);
        std::string s;
        while (std::getline(std::cin, s)) {
            if (s.empty()) continue;
            Node node = graph.buildNode(s.c_str());
            printf("%d %d\n", (int)graph.indegree(node), (int)graph.outdegree(node));
        }
    } catch (Exception& e) { fprintf(stderr, "EXCEPTION: %s\n", e.getMessage()); return 1; }
    return 0;
}
