"""Cell-type lineage container with the attribute surface of the reference's ``tree.py``
(``Tree.head()``, ``tree[id]``, ``add_head``, ``add_node``, ``display``, ``traverse``; nodes carry ``identifier``,
``children``, ``cellTypeMembers``, ``childCellTypeMembers``, ``proj``, ``const``,
``cellTypeMean``; reference tree.py:4-105).  Host-side data structure only.
"""
from collections import deque
from dataclasses import dataclass, field
from typing import Any, Dict, List

_ROOT, _DEPTH, _BREADTH = 0, 1, 2          # tree.py:1 (depth of the root for display; the two traversal modes)


@dataclass
class Node:
    identifier: int
    cellTypeMembers: Any
    const: Any
    proj: Any
    cellTypeMean: Any = 0
    children: List[int] = field(default_factory=list)
    childCellTypeMembers: Any = field(default_factory=list)

    def add_child(self, identifier):
        self.children.append(identifier)

    def setChildCellTypeMembers(self, members):
        self.childCellTypeMembers = members


class Tree:
    def __init__(self):
        self.nodes: Dict[int, Node] = {}
        self._root = None

    def head(self):
        return self._root

    def add_head(self, identifier, cellTypeMembers, const, proj, cellTypeMean):
        self._root = self.nodes[identifier] = Node(identifier, cellTypeMembers, const, proj, cellTypeMean)
        return self._root

    def add_node(self, identifier, cellTypeMembers, const, proj, cellTypeMean, parent):
        node = self.nodes[identifier] = Node(identifier, cellTypeMembers, const, proj, cellTypeMean)
        self.nodes[parent].add_child(identifier)
        return node

    def __getitem__(self, key):
        return self.nodes[key]

    def __setitem__(self, key, node):
        self.nodes[key] = node

    def walk(self, start=None):
        """Nodes in the depth-first order findAllNextStates visits them (MS:607-612)."""
        node = self._root if start is None else start
        yield node
        for child in node.children:
            yield from self.walk(self.nodes[child])

    def display(self, identifier=None, depth=_ROOT):
        """tree.py:56-65: the subtree below ``identifier`` (default: the head), one identifier per line, indented by depth."""
        node = self._root if identifier is None else self.nodes[identifier]
        if depth == _ROOT:
            print("{0}".format(node.identifier))
        else:
            print("\t" * depth, "{0}".format(node.identifier))
        for child in node.children:
            self.display(child, depth + 1)

    def describe(self, identifier):
        """The six description lines tree.py:71-79 reports per cell type."""
        node = self.nodes[identifier]
        return ["\ncellType Number: " + str(identifier), "cellType Const: " + str(node.const), "cellType Proj: " + str(node.proj),
                "cellType Mean: " + str(node.cellTypeMean), "Num cellType Members: " + str(len(node.cellTypeMembers)),
                "Num cellType Child Members: " + str(len(node.childCellTypeMembers))]

    def traverse(self, identifier, mode=_DEPTH):
        """tree.py:67-99: generator over the descriptions of ``identifier`` and everything below it, depth first (children
        before siblings, the order findAllNextStates steps the types in) or breadth first (``mode=_BREADTH``)."""
        yield self.describe(identifier)
        pending = deque(self.nodes[identifier].children)
        while pending:
            current = pending.popleft()
            yield self.describe(current)
            below = self.nodes[current].children
            if mode == _DEPTH:
                pending.extendleft(reversed(below))
            elif mode == _BREADTH:
                pending.extend(below)
