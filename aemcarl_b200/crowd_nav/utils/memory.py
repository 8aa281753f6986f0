"""Host-side replay memory with the interface the reference's callers use (crowd_nav/utils/memory.py:4-28):
`ReplayMemory(capacity)`, `push(item)`, `is_full()`, `len()`, indexing (a torch `Dataset` for the DataLoader of
trainer.py:61,70), `clear()`, and the public attributes `capacity`, `memory`, `position`.

Only the serial compatibility path (crowd_nav/utils/explorer.py, one episode at a time) stores experience here; the
vectorised path keeps it in HBM (`DeviceReplayMemory`, utils/vec_explorer.py), whose ring is tested against this one.
Behaviour kept from the reference: the write cursor `position` advances modulo `capacity` on every push and is NOT reset
by `clear()`; while fewer than `position + 1` items are stored a push appends, otherwise it overwrites the slot under the
cursor.
"""
from torch.utils.data import Dataset


class ReplayMemory(Dataset):
    def __init__(self, capacity):
        if int(capacity) < 1:
            raise ValueError("capacity must be positive")
        self.capacity = int(capacity)
        self.position = 0          # slot the next overwrite goes to
        self._items = []

    # the reference exposes the backing list; explorer / tests read it
    @property
    def memory(self):
        return self._items

    def push(self, item):
        slot = self.position
        self.position = (slot + 1) % self.capacity
        if slot < len(self._items):
            self._items[slot] = item           # ring is (partly) populated under the cursor: replace the oldest entry there
        else:
            self._items.append(item)           # still growing (also right after clear(), whatever the cursor says)

    def extend(self, items):
        for it in items:
            self.push(it)

    def is_full(self):
        return len(self._items) >= self.capacity

    def clear(self):
        self._items = []

    def __len__(self):
        return len(self._items)

    def __getitem__(self, index):
        return self._items[index]
