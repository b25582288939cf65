"""`Task`: what the optimizer needs to know about a unit of work.

The placement-relevant part of sky/task.py: `num_nodes` (:1076-1086),
inputs / outputs with their sizes (:1238-1274), `set_resources` accepting one
Resources, a set ("any of") or a list ("ordered") (:1292-1328), the time
estimator (:1361-1379) and `>>` (:2018-2024).
"""
from typing import Callable, List, Optional, Set, Union

from skypilot_b200 import clouds
from skypilot_b200 import dag as dag_lib
from skypilot_b200 import resources as resources_lib


class Task:

    def __init__(self, name: Optional[str] = None, *,
                 setup: Optional[str] = None, run: Optional[str] = None,
                 num_nodes: Optional[int] = None):
        self.name = name
        self.setup = setup
        self.run = run
        self._num_nodes = 1
        self.num_nodes = num_nodes
        self.inputs: Optional[str] = None
        self.outputs: Optional[str] = None
        self.estimated_inputs_size_gigabytes: Optional[float] = None
        self.estimated_outputs_size_gigabytes: Optional[float] = None
        self.time_estimator_func: Optional[Callable[
            ['resources_lib.Resources'], int]] = None
        self.resources: Union[List[resources_lib.Resources],
                              Set[resources_lib.Resources]] = {
                                  resources_lib.Resources()
                              }
        self.best_resources: Optional[resources_lib.Resources] = None
        dag = dag_lib.get_current_dag()
        if dag is not None:
            dag.add(self)

    @property
    def num_nodes(self) -> int:
        return self._num_nodes

    @num_nodes.setter
    def num_nodes(self, num_nodes: Optional[int]) -> None:
        if num_nodes is None:
            num_nodes = 1
        if not isinstance(num_nodes, int) or num_nodes <= 0:
            raise ValueError(
                f'num_nodes should be a positive int. Got: {num_nodes}')
        self._num_nodes = num_nodes

    def set_inputs(self, inputs: str, estimated_size_gigabytes: float) -> 'Task':
        self.inputs = inputs
        self.estimated_inputs_size_gigabytes = estimated_size_gigabytes
        return self

    def get_inputs(self) -> Optional[str]:
        return self.inputs

    def get_estimated_inputs_size_gigabytes(self) -> Optional[float]:
        return self.estimated_inputs_size_gigabytes

    def get_inputs_cloud(self):
        """The cloud the inputs live in, from the URL scheme."""
        assert isinstance(self.inputs, str), self.inputs
        if self.inputs.startswith('s3:'):
            return clouds.AWS()
        if self.inputs.startswith('gs:'):
            return clouds.GCP()
        raise ValueError(f'cloud path not supported: {self.inputs}')

    def set_outputs(self, outputs: str,
                    estimated_size_gigabytes: float) -> 'Task':
        self.outputs = outputs
        self.estimated_outputs_size_gigabytes = estimated_size_gigabytes
        return self

    def get_outputs(self) -> Optional[str]:
        return self.outputs

    def get_estimated_outputs_size_gigabytes(self) -> Optional[float]:
        return self.estimated_outputs_size_gigabytes

    def set_resources(self, resources) -> 'Task':
        """One Resources, a set (any of, unordered) or a list (ordered)."""
        if isinstance(resources, resources_lib.Resources):
            resources = {resources}
        self.resources = resources
        return self

    def set_resources_override(self, override_params) -> 'Task':
        """Applies the overrides to every alternative
        (sky/task.py:1330-1338)."""
        new = [res.copy(**override_params) for res in list(self.resources)]
        self.set_resources(type(self.resources)(new))
        return self

    def set_time_estimator(self, func) -> 'Task':
        self.time_estimator_func = func
        return self

    def estimate_runtime(self, resources):
        if self.time_estimator_func is None:
            raise NotImplementedError(
                f'Node [{self}] does not have a cost model set; '
                'call set_time_estimator() first')
        return self.time_estimator_func(resources)

    def __rshift__(self, other: 'Task') -> 'Task':
        dag_lib.get_current_dag().add_edge(self, other)
        return other

    def __repr__(self) -> str:
        if self.name is not None:
            return f'Task({self.name})'
        return f'Task(run={self.run!r})' if self.run else 'Task(<unnamed>)'
