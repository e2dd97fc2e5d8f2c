"""Warehouse layouts: grid size, highway bitmap and goal cells.

Host-side construction-time code (cold path).  Mirrors the reference's two layout builders:
`Warehouse._make_layout_from_params` (rware/warehouse.py:294-326) and `_make_layout_from_str`
(rware/warehouse.py:328-350).
"""
import numpy as np


def layout_from_params(shelf_columns: int, shelf_rows: int, column_height: int):
    """Returns (grid_size=(H, W), highways uint8[H, W], goals [(x, y), ...]).  rware/warehouse.py:294-326."""
    assert shelf_columns % 2 == 1, "Only odd number of shelf columns is supported"
    H = (column_height + 1) * shelf_rows + 2
    W = (2 + 1) * shelf_columns + 1
    goals = [(W // 2 - 1, H - 1), (W // 2, H - 1)]
    x = np.arange(W)[None, :]
    y = np.arange(H)[:, None]
    highways = (
        (x % 3 == 0)                                # vertical corridors
        | (y % (column_height + 1) == 0)            # horizontal corridors
        | (y == H - 1)                              # delivery row
        | ((y > H - (column_height + 3)) & ((x == W // 2 - 1) | (x == W // 2)))  # queue in front of the goals
    )
    return (H, W), np.ascontiguousarray(highways, dtype=np.uint8), goals


def layout_from_str(layout: str):
    """'x' shelf cell, '.' corridor, 'g' goal (also a corridor).  rware/warehouse.py:328-350."""
    layout = layout.strip().replace(" ", "")
    lines = layout.split("\n")
    H, W = len(lines), len(lines[0])
    for line in lines:
        assert len(line) == W, "Layout must be rectangular"
    highways = np.zeros((H, W), dtype=np.uint8)
    goals = []
    for y, line in enumerate(lines):
        for x, char in enumerate(line):
            assert char.lower() in "gx."
            if char.lower() == "g":
                goals.append((x, y))
                highways[y, x] = 1
            elif char.lower() == ".":
                highways[y, x] = 1
    assert len(goals) >= 1, "At least one goal is required"
    return (H, W), highways, goals
