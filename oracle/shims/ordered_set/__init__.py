"""Shim for ordered_set.OrderedSet (insertion-ordered set). TEST ONLY. data_loader.py:11,64."""


class OrderedSet(dict):
    def add(self, key):
        self.setdefault(key, len(self))

    def __iter__(self):
        return iter(self.keys())
