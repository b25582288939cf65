"""`Dag`: tasks plus dependency edges (the optimizer-facing part of
sky/dag.py: context manager, add / remove, `>>` edges, `is_chain` :159-178)."""
import enum
import threading
from typing import List, Optional

import networkx as nx


class DagExecution(enum.Enum):
    """How the tasks of a multi-task DAG run (sky/dag.py:12-19)."""
    SERIAL = 'serial'      # pipeline
    PARALLEL = 'parallel'  # job group: all tasks start together


class Dag:
    """A directed acyclic graph of Tasks; `with Dag() as dag:` collects the
    tasks created inside the block."""

    def __init__(self) -> None:
        self.tasks: List['object'] = []
        self.graph = nx.DiGraph()
        self.name: Optional[str] = None
        self.execution: Optional[DagExecution] = None

    def is_job_group(self) -> bool:
        """Parallel execution mode makes a DAG a JobGroup (sky/dag.py:91-97)."""
        return self.execution == DagExecution.PARALLEL

    def set_execution(self, execution: DagExecution) -> None:
        self.execution = execution

    def add(self, task) -> None:
        self.graph.add_node(task)
        self.tasks.append(task)

    def remove(self, task) -> None:
        self.tasks.remove(task)
        self.graph.remove_node(task)

    def add_edge(self, op1, op2) -> None:
        assert op1 in self.graph.nodes
        assert op2 in self.graph.nodes
        self.graph.add_edge(op1, op2)

    def __len__(self) -> int:
        return len(self.tasks)

    def __enter__(self) -> 'Dag':
        push_dag(self)
        return self

    def __exit__(self, exc_type, exc_value, traceback) -> None:
        pop_dag()

    def __repr__(self) -> str:
        return f'DAG:\n ' + '\n '.join(repr(t) for t in self.tasks)

    def get_graph(self):
        return self.graph

    def is_chain(self) -> bool:
        """True iff the tasks form one linear chain: every node has at most
        one parent / child and exactly one node has none."""
        succ, pred = self.graph._succ, self.graph._pred  # pylint: disable=protected-access
        if not succ:
            return True
        # (the adjacency dicts directly: a degree view per node is most of
        # what this costs on a small DAG)
        roots = leaves = 0
        for n, out in succ.items():
            if len(out) > 1 or len(pred[n]) > 1:
                return False
            leaves += not out
            roots += not pred[n]
        return roots == 1 and leaves == 1


class _DagContext(threading.local):
    """Per-thread stack of the DAGs being built."""

    def __init__(self):
        super().__init__()
        self.current: Optional[Dag] = None
        self.stack: List[Dag] = []

    def push(self, dag: Dag) -> None:
        self.stack.append(dag)
        self.current = dag

    def pop(self) -> Optional[Dag]:
        old = self.stack.pop()
        self.current = self.stack[-1] if self.stack else None
        return old


_context = _DagContext()
push_dag = _context.push
pop_dag = _context.pop


def get_current_dag() -> Optional[Dag]:
    return _context.current
