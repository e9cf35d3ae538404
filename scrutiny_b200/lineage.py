"""Cell-type lineage container with the attribute surface of the reference's ``tree.py``
(``Tree.head()``, ``tree[id]``, ``add_head``, ``add_node``; nodes carry ``identifier``,
``children``, ``cellTypeMembers``, ``childCellTypeMembers``, ``proj``, ``const``,
``cellTypeMean``; reference tree.py:4-54).  Host-side data structure only.
"""
from dataclasses import dataclass, field
from typing import Any, Dict, List


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

    def display(self, identifier=None, depth=0):
        node = self._root if identifier is None else self.nodes[identifier]
        print("\t" * depth + str(node.identifier))
        for child in node.children:
            self.display(child, depth + 1)
