// Host side of the per-cell <data type="time_course"> likelihood: the assignment of observed to simulated cells.
//
// The reference matches with dependencies/hungarian2/hungarian.cpp (James Payor's alternating-tree Hungarian algorithm on an
// edge list, as the reference's authors modified it: n_right / edge_count arguments, and the initial partition into "tight" edges
// tests an INT -- `int reducedCost = cost - potentials; if (reducedCost < 1e-12)`, hungarian.cpp:170-172 -- so every edge whose
// reduced cost lies below 1 starts out as tight). The greedy initial matching over those edges is kept by the search that
// follows, and the result is in general NOT the minimum-cost matching: on the golden fixtures the matched log-likelihood is
// 0.3-5 units below the optimum. DataLikelihoodTimeCourse::Evaluate (.cpp:323-336) sums the likelihoods of exactly that
// matching, so the matching is part of the likelihood's value and is restated here step for step -- the same edge arrays with
// the same swaps (they decide the order in which the search visits the edges), the same comparisons in double, the same
// first-found / strictly-smaller choices -- for the only case the likelihood produces: a complete cost matrix with as many
// observed (left) as simulated (right) cells. Checked against the reference's own compiled implementation on the golden
// fixtures (the compiled-reference checker links it) and on random matrices (tests/test_cellpop_cpu.py).
//
// One addition: a bound on the number of trips of the search (the reference has none; with potentials that start out
// inconsistent there is no proof of termination). Exceeding it returns "no matching", i.e. a log-likelihood of -inf.
#pragma once

#include <cfloat>
#include <deque>
#include <utility>
#include <vector>

namespace bcm3b200 {

// cost: [n][n] row-major, row = observed cell, column = simulated cell. Returns the simulated cell matched to every observed
// cell, or an empty vector when no perfect matching was found.
inline std::vector<int> payor_matching_complete(int n, const double* cost)
{
	struct Edge {
		int right;
		double cost;
	};
	const double oo = DBL_MAX;
	const int NONE = -1;
	if (n <= 0) return std::vector<int>();
	// the sorted, duplicate-free edge list of every left node: all n right nodes in order
	std::vector<Edge> edges((size_t)n * n);
	for (int i = 0; i < n; i++)
		for (int j = 0; j < n; j++) edges[(size_t)i * n + j] = Edge{ j, cost[(size_t)i * n + j] };
	auto E = [&](int i) { return edges.data() + (size_t)i * n; };

	// potentials: left = the cheapest edge of the node; right = the cheapest reduced cost that reaches the node
	std::vector<double> left_pot(n), right_pot(n, oo);
	for (int i = 0; i < n; i++) {
		double smallest = E(i)[0].cost;
		for (int e = 1; e < n; e++)
			if (E(i)[e].cost < smallest) smallest = E(i)[e].cost;
		left_pot[i] = smallest;
	}
	for (int i = 0; i < n; i++)
		for (int j = 0; j < n; j++) {
			const double reduced = cost[(size_t)i * n + j] - left_pot[i];
			if (right_pot[j] > reduced) right_pot[j] = reduced;
		}

	// tight edges to the front of every list -- with the reference's integer test (see the header)
	std::vector<int> tight(n, 0);
	for (int i = 0; i < n; i++) {
		int count = 0;
		for (int e = 0; e < n; e++) {
			const int reduced = (int)(E(i)[e].cost - left_pot[i] - right_pot[E(i)[e].right]);
			if (reduced < 1e-12) {
				if (e != count) std::swap(E(i)[count], E(i)[e]);
				++count;
			}
		}
		tight[i] = count;
	}

	// greedy initial matching over the tight edges
	int cardinality = 0;
	std::vector<int> left_match(n, NONE), right_match(n, NONE);
	for (int i = 0; i < n; i++)
		for (int e = 0; e < tight[i]; e++) {
			const int j = E(i)[e].right;
			if (right_match[j] == NONE) {
				++cardinality;
				right_match[j] = i;
				left_match[i] = j;
				break;
			}
		}
	if (cardinality == n) return left_match;

	std::vector<double> min_slack(n);
	std::vector<int> min_slack_left(n), min_slack_edge(n), backtrack(n);
	std::vector<char> seen(n);
	std::deque<int> queue;
	long long trips = 0;
	const long long trip_limit = 64ll * n * n + 4096;

	auto unreached = [&](int j) { return right_match[j] == NONE || !seen[right_match[j]]; };
	while (cardinality < n) {
		for (int j = 0; j < n; j++) {
			min_slack[j] = oo;
			min_slack_left[j] = NONE;
			backtrack[j] = NONE;
			seen[j] = 0;
		}
		queue.clear();
		// start from the unmatched left node with the fewest tight edges (the first of them)
		int start = NONE;
		{
			double fewest = oo;
			for (int i = 0; i < n; i++)
				if (left_match[i] == NONE && tight[i] < fewest) {
					fewest = tight[i];
					start = i;
				}
		}
		queue.push_back(start);
		seen[start] = 1;
		int end = NONE;
		// following an edge (i, j) of the tight subgraph: a free right node ends the search, a matched one extends the tree
		auto follow = [&](int i, int j) {
			backtrack[j] = i;
			const int m = right_match[j];
			if (m == NONE) {
				end = j;
			} else if (!seen[m]) {
				seen[m] = 1;
				queue.push_back(m);
			}
		};
		while (end == NONE) {
			if (++trips > trip_limit) return std::vector<int>();
			while (end == NONE && !queue.empty()) {
				const int i = queue.front();
				queue.pop_front();
				Edge* ed = E(i);
				for (int e = 0; e < tight[i]; ++e) { // every tight edge is visited, also after a free node was found
					const int j = ed[e].right;
					if (ed[e].cost > left_pot[i] + right_pot[j]) { // not tight any more: behind the tight ones
						--tight[i];
						std::swap(ed[e], ed[tight[i]]);
						--e;
						continue;
					}
					if (backtrack[j] != NONE) continue;
					follow(i, j);
				}
				if (end == NONE) {
					const double pot = left_pot[i];
					for (int e = tight[i]; e < n; e++) {
						const int j = ed[e].right;
						if (unreached(j)) {
							const double reduced = ed[e].cost - pot - right_pot[j];
							if (reduced < min_slack[j]) {
								min_slack[j] = reduced;
								min_slack_left[j] = i;
								min_slack_edge[j] = e;
							}
						}
					}
				}
			}
			if (end == NONE) {
				// the tree cannot grow: shift the potentials by the smallest slack towards an unreached right node
				int min_j = NONE;
				double slack = oo;
				for (int j = 0; j < n; j++)
					if (unreached(j) && min_slack[j] < slack) {
						slack = min_slack[j];
						min_j = j;
					}
				if (min_j == NONE || min_slack_left[min_j] == NONE) return std::vector<int>();
				for (int i = 0; i < n; i++)
					if (seen[i]) {
						left_pot[i] += slack;
						if (left_match[i] != NONE) right_pot[left_match[i]] -= slack;
					}
				for (int j = 0; j < n; j++)
					if (unreached(j)) {
						min_slack[j] -= slack;
						if (min_slack[j] == 0) { // a new tight edge
							const int i = min_slack_left[j], e = min_slack_edge[j];
							if (e != tight[i]) std::swap(E(i)[e], E(i)[tight[i]]);
							++tight[i];
							if (end == NONE) follow(i, j);
						}
					}
			}
		}
		++cardinality;
		for (int cur = end; cur != NONE;) { // flip the augmenting path
			const int l = backtrack[cur], next = left_match[l];
			right_match[cur] = l;
			left_match[l] = cur;
			cur = next;
		}
	}
	return left_match;
}

} // namespace bcm3b200
