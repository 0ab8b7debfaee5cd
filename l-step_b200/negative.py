"""Drop-in for the reference's NegativeEdgeSampler (utils/utils.py:304-494) — SURVEY §8(f) f3.

Same constructor, attributes, methods and — for a fixed seed — bit-identical samples for all three strategies
('random', 'historical', 'inductive'). What is different is the cost per batch of the two strategies BASELINE config 3 evaluates:

  * the reference rebuilds `historical_edges` = the Python set of ALL (src, dst) pairs seen up to the batch's start time from
    scratch in every call (a boolean mask over every edge of the graph plus up to E tuple insertions: ~0.3-0.5 s per
    batch at Reddit size), and materialises `possible_edges` = |unique src| x |unique dst| tuples (9.84 M for the
    Reddit shape, ~10 s and ~2 GB) in the constructor although only the fall-back branch of the very first batches reads it;
  * here the history set is kept between calls and only the edges that became visible since the previous call are
    inserted (edge streams are chronological: utils/DataLoader.py:199,229-242), and `possible_edges` is built on first use.

Why that is still RNG-bit-identical: the samples depend on numpy's RandomState stream (same calls, same arguments, same
order) and on the ITERATION ORDER of CPython sets (`np.array([edge[0] for edge in unique_historical_edges])`,
`list(self.possible_edges - batch_edges)`). A CPython set's slot layout is a deterministic function of the sequence of
insertions (hash values and order) — `set(generator)` performs exactly the insertions, one by one, that incremental
`set.update` performs on the same sequence — and every derived set (`a - b`) is produced by the interpreter's own set
code from identically laid out operands. So the incremental set is, slot for slot, the set the reference would rebuild,
and everything downstream of it is the reference's own expression evaluated on equal inputs. When a call goes back in
time (a new evaluation pass after reset_random_state) or the stream is not chronological, the history is rebuilt the
reference's way. This is host code by nature ("likely an incremental host structure, not CUDA").
"""
from __future__ import annotations

from itertools import chain as _chain

import numpy as np


class NegativeEdgeSampler(object):

    def __init__(self, src_node_ids: np.ndarray, dst_node_ids: np.ndarray, interact_times: np.ndarray = None, last_observed_time: float = None,
                 negative_sample_strategy: str = 'random', seed: int = None):
        self.seed = seed
        self.negative_sample_strategy = negative_sample_strategy
        self.src_node_ids = src_node_ids
        self.dst_node_ids = dst_node_ids
        self.interact_times = interact_times
        self.unique_src_node_ids = np.unique(src_node_ids)
        self.unique_dst_node_ids = np.unique(dst_node_ids)
        self.unique_interact_times = np.unique(interact_times)  # (np.unique(None) -> array([None]), as in the reference)
        self.earliest_time = min(self.unique_interact_times)
        self.last_observed_time = last_observed_time
        self._possible_edges = None  # built on first use (utils/utils.py:328-330 builds it eagerly)
        self._hist = None            # incremental historical_edges: set, number of stream edges inserted, its end time
        self._hist_n = 0
        self._hist_end = None
        self._chronological = None
        if self.negative_sample_strategy == 'inductive':
            # set of observed edges (utils/utils.py:332-334)
            self.observed_edges = self.get_unique_edges_between_start_end_time(self.earliest_time, self.last_observed_time)
        if self.seed is not None:
            self.random_state = np.random.RandomState(self.seed)

    # ---- the reference's full-scan form (kept: public method, and the fall-back for non-chronological use) -------------
    def get_unique_edges_between_start_end_time(self, start_time: float, end_time: float):
        selected_time_interval = np.logical_and(self.interact_times >= start_time, self.interact_times <= end_time)
        return set((src_node_id, dst_node_id) for src_node_id, dst_node_id in
                   zip(self.src_node_ids[selected_time_interval], self.dst_node_ids[selected_time_interval]))

    @property
    def possible_edges(self):
        """All |unique src| x |unique dst| pairs, in the reference's construction order (utils/utils.py:328-330)."""
        if self._possible_edges is None:
            self._possible_edges = set((src_node_id, dst_node_id) for src_node_id in self.unique_src_node_ids
                                       for dst_node_id in self.unique_dst_node_ids)
        return self._possible_edges

    def _historical_edges(self, end_time: float):
        """== get_unique_edges_between_start_end_time(self.earliest_time, end_time), slot for slot, maintained incrementally."""
        t = self.interact_times
        if self._chronological is None:
            self._chronological = bool(len(t) == 0 or np.all(t[1:] >= t[:-1]))
        if not self._chronological:
            return self.get_unique_edges_between_start_end_time(self.earliest_time, end_time)
        n = int(np.searchsorted(t, end_time, side='right'))  # edges with earliest <= t <= end_time: a prefix of the stream
        if self._hist is None or n < self._hist_n:  # first call, or a call that goes back in time: rebuild
            self._hist, self._hist_n = set(), 0
        if n > self._hist_n:
            lo = self._hist_n
            # the same insertions, in the same order, that set(generator over the prefix) performs
            self._hist.update(zip(self.src_node_ids[lo:n], self.dst_node_ids[lo:n]))
            self._hist_n = n
        self._hist_end = end_time
        return self._hist

    # ---- sampling (utils/utils.py:350-487) ------------------------------------------------------------------------------
    def sample(self, size: int, batch_src_node_ids: np.ndarray = None, batch_dst_node_ids: np.ndarray = None,
               current_batch_start_time: float = 0.0, current_batch_end_time: float = 0.0):
        if self.negative_sample_strategy == 'random':
            negative_src_node_ids, negative_dst_node_ids = self.random_sample(size=size)
        elif self.negative_sample_strategy == 'historical':
            negative_src_node_ids, negative_dst_node_ids = self.historical_sample(size=size, batch_src_node_ids=batch_src_node_ids,
                                                                                  batch_dst_node_ids=batch_dst_node_ids,
                                                                                  current_batch_start_time=current_batch_start_time,
                                                                                  current_batch_end_time=current_batch_end_time)
        elif self.negative_sample_strategy == 'inductive':
            negative_src_node_ids, negative_dst_node_ids = self.inductive_sample(size=size, batch_src_node_ids=batch_src_node_ids,
                                                                                 batch_dst_node_ids=batch_dst_node_ids,
                                                                                 current_batch_start_time=current_batch_start_time,
                                                                                 current_batch_end_time=current_batch_end_time)
        else:
            raise ValueError(f'Not implemented error for negative_sample_strategy {self.negative_sample_strategy}!')
        return negative_src_node_ids, negative_dst_node_ids

    def random_sample(self, size: int):
        if self.seed is None:
            random_sample_edge_src_node_indices = np.random.randint(0, len(self.unique_src_node_ids), size)
            random_sample_edge_dst_node_indices = np.random.randint(0, len(self.unique_dst_node_ids), size)
        else:
            random_sample_edge_src_node_indices = self.random_state.randint(0, len(self.unique_src_node_ids), size)
            random_sample_edge_dst_node_indices = self.random_state.randint(0, len(self.unique_dst_node_ids), size)
        return self.unique_src_node_ids[random_sample_edge_src_node_indices], self.unique_dst_node_ids[random_sample_edge_dst_node_indices]

    def random_sample_with_collision_check(self, size: int, batch_src_node_ids: np.ndarray, batch_dst_node_ids: np.ndarray):
        assert batch_src_node_ids is not None and batch_dst_node_ids is not None
        batch_edges = set((batch_src_node_id, batch_dst_node_id) for batch_src_node_id, batch_dst_node_id in zip(batch_src_node_ids, batch_dst_node_ids))
        possible_random_edges = list(self.possible_edges - batch_edges)
        assert len(possible_random_edges) > 0
        random_edge_indices = self.random_state.choice(len(possible_random_edges), size=size, replace=len(possible_random_edges) < size)
        return np.array([possible_random_edges[random_edge_idx][0] for random_edge_idx in random_edge_indices]), \
               np.array([possible_random_edges[random_edge_idx][1] for random_edge_idx in random_edge_indices])

    def _from_edge_set(self, size, edges, batch_src_node_ids, batch_dst_node_ids):
        """Common tail of historical_sample / inductive_sample (utils/utils.py:424-441, 465-482)."""
        if len(edges):  # one pass over the set in its iteration order (== the reference's two list comprehensions)
            pairs = np.fromiter(_chain.from_iterable(edges), dtype=np.int64, count=2 * len(edges)).reshape(-1, 2)
            edges_src_node_ids, edges_dst_node_ids = pairs[:, 0], pairs[:, 1]
        else:
            edges_src_node_ids = edges_dst_node_ids = np.array([])
        if size > len(edges):
            num_random_sample_edges = size - len(edges)
            random_sample_src_node_ids, random_sample_dst_node_ids = self.random_sample_with_collision_check(
                size=num_random_sample_edges, batch_src_node_ids=batch_src_node_ids, batch_dst_node_ids=batch_dst_node_ids)
            negative_src_node_ids = np.concatenate([random_sample_src_node_ids, edges_src_node_ids])
            negative_dst_node_ids = np.concatenate([random_sample_dst_node_ids, edges_dst_node_ids])
        else:
            sample_edge_node_indices = self.random_state.choice(len(edges), size=size, replace=False)
            negative_src_node_ids = edges_src_node_ids[sample_edge_node_indices]
            negative_dst_node_ids = edges_dst_node_ids[sample_edge_node_indices]
        # (if one input of np.concatenate is empty the output is float: convert, as the reference does)
        return negative_src_node_ids.astype(np.longlong), negative_dst_node_ids.astype(np.longlong)

    def historical_sample(self, size: int, batch_src_node_ids: np.ndarray, batch_dst_node_ids: np.ndarray,
                          current_batch_start_time: float, current_batch_end_time: float):
        assert self.seed is not None
        historical_edges = self._historical_edges(current_batch_start_time)
        current_batch_edges = self.get_unique_edges_between_start_end_time(start_time=current_batch_start_time, end_time=current_batch_end_time) \
            if not self._chronological else self._window_edges(current_batch_start_time, current_batch_end_time)
        unique_historical_edges = historical_edges - current_batch_edges
        return self._from_edge_set(size, unique_historical_edges, batch_src_node_ids, batch_dst_node_ids)

    def inductive_sample(self, size: int, batch_src_node_ids: np.ndarray, batch_dst_node_ids: np.ndarray,
                         current_batch_start_time: float, current_batch_end_time: float):
        assert self.seed is not None
        historical_edges = self._historical_edges(current_batch_start_time)
        current_batch_edges = self.get_unique_edges_between_start_end_time(start_time=current_batch_start_time, end_time=current_batch_end_time) \
            if not self._chronological else self._window_edges(current_batch_start_time, current_batch_end_time)
        unique_inductive_edges = historical_edges - self.observed_edges - current_batch_edges
        return self._from_edge_set(size, unique_inductive_edges, batch_src_node_ids, batch_dst_node_ids)

    def _window_edges(self, start_time: float, end_time: float):
        """== get_unique_edges_between_start_end_time(start_time, end_time) on a chronological stream: the same slice, found by
        two binary searches instead of a mask over every edge."""
        t = self.interact_times
        lo = int(np.searchsorted(t, start_time, side='left'))
        hi = int(np.searchsorted(t, end_time, side='right'))
        return set((src_node_id, dst_node_id) for src_node_id, dst_node_id in zip(self.src_node_ids[lo:hi], self.dst_node_ids[lo:hi]))

    def reset_random_state(self):
        self.random_state = np.random.RandomState(self.seed)
