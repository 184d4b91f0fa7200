"""2-D sin-cos relative position table (reference: encoder/gcn_lib/pos_embed.py).

Only used at construction time to fill Grapher's frozen ``relative_pos`` parameter, which the
forward pass never reads (torch_vertex.py:188-190); kept so ``state_dict`` shapes and values
match the reference.
"""
import numpy as np


def get_1d_sincos_pos_embed_from_grid(embed_dim, pos):
    """pos: (M,) positions -> (M, embed_dim) = [sin(pos * w) | cos(pos * w)], w_i = 10000^(-2i/embed_dim)."""
    assert embed_dim % 2 == 0
    freq = 1.0 / 10000 ** (np.arange(embed_dim // 2, dtype=np.float64) / (embed_dim / 2.0))
    phase = np.outer(np.reshape(pos, -1), freq)
    return np.concatenate([np.sin(phase), np.cos(phase)], axis=1)


def get_2d_sincos_pos_embed_from_grid(embed_dim, grid):
    assert embed_dim % 2 == 0
    halves = [get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[i]) for i in (0, 1)]
    return np.concatenate(halves, axis=1)


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    """(grid_size**2 [+1], embed_dim) embedding of a square grid; the w coordinate varies fastest."""
    axis = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(axis, axis), axis=0).reshape(2, 1, grid_size, grid_size)
    emb = get_2d_sincos_pos_embed_from_grid(embed_dim, grid)
    if cls_token:
        emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return emb


def get_2d_relative_pos_embed(embed_dim, grid_size):
    """(grid_size**2, grid_size**2) table 2 * E E^T / embed_dim."""
    emb = get_2d_sincos_pos_embed(embed_dim, grid_size)
    return 2 * np.matmul(emb, emb.transpose()) / emb.shape[1]
