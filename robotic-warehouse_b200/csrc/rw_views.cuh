// rw_views.cuh -- whole-warehouse views computed from the device state:
//   * get_global_image (rware/warehouse.py:966-1040): stacked float32 layers of the full H x W grid, zero-padded to pad_to_shape
//   * the rgb_array frame (rware/rendering.py:121-333): the Viewer's grid / goal / shelf / agent geometry rasterised per pixel
// Both are gathers: one block stages the 32 worlds of a tile, builds a one-byte occupancy map per world and then every
// thread computes output elements from it, so global stores are coalesced over the output tensor.  The per-element
// functions are host/device so that the test-only host twin (tests/hostsim) runs the same source.
#pragma once
#include "rw_common.h"

struct RwViewSpec {      // get_global_image arguments, resolved on the host
    int32_t n_layers, layers[8];  // ImageLayer values (warehouse.py:59-70)
    int32_t C, Hp, Wp;            // output shape per world (after padding)
    int32_t c0, y0, x0;           // pad_before per axis (warehouse.py:1028: pad_dim // 2)
};

// Occupancy byte per cell as _recalc_grid (749-755) and the layer loops see the world: RW_G_AGENT bits = id of the LAST agent
// written to the cell, RW_G_SHELF = some shelf stands there (grid[_LAYER_SHELFS] > 0), RW_G_REQ = some requested shelf
// stands there (992-995).  `l.g` must be zeroed.  Returns how many index errors the reference would raise for the transposed
// AGENT_DIRECTION / AGENT_LOAD writes `layer[ag.x, ag.y]` (1004, 1009) -- at most one per world is reported.
RW_HD int rw_view_build(const RwParams &p, const RwLane &l, bool has_dir_layer, bool has_load_layer) {
    const RwOff o = rw_off<RwDynamic>(p);
    for (int k = 0; k < p.S; k++) l.g[rw_shelf_cell(o, l, k)] |= (uint8_t)RW_G_SHELF;
    for (int q = 0; q < p.R; q++) {
        int id = rw_queue(o, l, q);
        if (id > 0) l.g[rw_shelf_cell(o, l, id - 1)] |= (uint8_t)RW_G_REQ;
    }
    int bad = 0;
    for (int i = 0; i < p.n; i++) {
        uint32_t aw = rw_agent(l, i);
        int x = rw_ax(aw), y = rw_ay(aw), c = y * p.W + x;
        l.g[c] = (uint8_t)((l.g[c] & (RW_G_SHELF | RW_G_REQ)) | (uint32_t)(i + 1));
        if ((x >= p.H || y >= p.W) && (has_dir_layer || (has_load_layer && rw_acarry(aw)))) bad = 1;
    }
    return bad;
}

// Every output element is "classified" once (which cell it looks at and how; independent of the world) and then "shaded" per
// world from that world's occupancy byte, so a thread that owns an element position pays the index arithmetic once per tile.
// code: bits 0-15 cell, 16-23 mask over the occupancy byte, 24 invert, 25-27 special (3 / 4: transposed agent layers)
#define RW_GC_INV (1u << 24)
#define RW_GC_SPECIAL(code) (((code) >> 25) & 7u)

// element [cp][yp][xp] of the padded output: np.pad constant 0 outside (1033-1038), else layer s.layers[c] at row y, column x
RW_HD uint32_t rw_gimg_classify(const RwParams &p, const uint8_t *goal_flags, const RwViewSpec &s, int cp, int yp, int xp) {
    const int c = cp - s.c0, y = yp - s.y0, x = xp - s.x0;
    if (c < 0 || c >= s.n_layers || y < 0 || y >= p.H || x < 0 || x >= p.W) return 0u;  // mask 0, no invert: constant 0
    const uint32_t cell = (uint32_t)(y * p.W + x);
    switch (s.layers[c]) {
        case 0: return cell | (RW_G_SHELF << 16);             // SHELVES 988-991
        case 1: return cell | (RW_G_REQ << 16);               // REQUESTS 992-995
        case 2: return cell | (RW_G_AGENT << 16);             // AGENTS 996-999
        case 6: return cell | (RW_G_AGENT << 16) | RW_GC_INV; // ACCESSIBLE 1014-1017
        case 5: return (goal_flags[cell] & 2) ? RW_GC_INV : 0u;  // GOALS 1010-1013 (the swapped loop names cancel: layer[y][x])
        case 3:
        case 4:  // layer[ag.x, ag.y] (sic, 1004 / 1009): element [y][x] belongs to the agent standing at column y, row x
            if (x >= p.H || y >= p.W) return 0u;
            return (uint32_t)(x * p.W + y) | ((uint32_t)s.layers[c] << 25);
    }
    return 0u;
}

RW_HD float rw_gimg_shade(const RwLane &l, uint32_t code) {
    const uint32_t gb = l.g[code & 0xffffu];
    const uint32_t special = RW_GC_SPECIAL(code);
    if (special) {
        const uint32_t id = gb & RW_G_AGENT;
        if (!id) return 0.0f;
        const uint32_t aw = rw_agent(l, (int)id - 1);
        if (special == 3) return (float)(rw_adir(aw) + 1);  // AGENT_DIRECTION 1000-1004
        return rw_acarry(aw) ? 1.0f : 0.0f;                 // AGENT_LOAD 1005-1009
    }
    const bool on = (gb & (code >> 16) & 0xffu) != 0;
    return (on != ((code & RW_GC_INV) != 0)) ? 1.0f : 0.0f;
}

RW_HD float rw_gimg_element(const RwParams &p, const RwLane &l, const uint8_t *goal_flags, const RwViewSpec &s, int idx) {
    const int plane = s.Hp * s.Wp;
    const int cp = idx / plane, rem = idx - cp * plane, yp = rem / s.Wp, xp = rem - yp * s.Wp;
    return rw_gimg_shade(l, rw_gimg_classify(p, goal_flags, s, cp, yp, xp));
}

// ---------------------------------------------------------------- rgb_array frame (rware/rendering.py)
// Viewer geometry (rendering.py:88-99): grid_size 30, window width 1 + W*31, height 2 + H*31; the frame is the colour buffer
// flipped to a top-left origin (136).  Draw order: white clear, black grid lines (140-178), grey goal cells (207-236), shelf
// quads with 2 px padding, teal when requested (180-205), agents as hexagons of radius 10, red when loaded (262-296), then a
// black heading line from the centre (298-333).  OpenGL is not available where this was written, so pixel coverage follows
// the usual rules instead of a recorded frame: a filled primitive covers the pixels whose centres it contains, an
// axis-aligned 1 px line at integer coordinate c covers the pixel row/column c - 1 (the Viewer offsets every line by +1 for
// exactly that reason: the lines of cell boundary k land on pixel 31 k and the window is 31 W + 1 wide).  The goal label
// glyph "G" (238-258) is a font rendering and is not drawn.
#define RW_RENDER_CELL 31
RW_HD int rw_render_width(const RwParams &p) { return 1 + p.W * RW_RENDER_CELL; }
RW_HD int rw_render_height(const RwParams &p) { return 2 + p.H * RW_RENDER_CELL; }

RW_HD uint32_t rw_rgb(uint32_t r, uint32_t g, uint32_t b) { return r | (g << 8) | (b << 16); }

// A frame pixel is classified once (constant colour, or a cell plus which primitives of that cell cover the pixel) and shaded
// per world.  code: bit 31 constant (bits 0-23 colour); else bits 0-15 cell, 16-19 covered by the heading line of direction d,
// 20 inside the agent hexagon, 21 inside the shelf quad, 22 goal cell.
RW_HD uint32_t rw_render_classify(const RwParams &p, const uint8_t *goal_flags, int i, int j) {
    const uint32_t white = rw_rgb(255, 255, 255) | (1u << 31), black = rw_rgb(0, 0, 0) | (1u << 31);
    const int gy = rw_render_height(p) - 1 - i;  // OpenGL row (origin bottom-left)
    const int cx = j / RW_RENDER_CELL, px = j - cx * RW_RENDER_CELL;  // px == 0: vertical grid line
    const int cyg = gy / RW_RENDER_CELL, py = gy - cyg * RW_RENDER_CELL;  // py == 0: horizontal grid line
    // grid lines: horizontal x in [0, 31 W), rows 31 r (r = 0..H); vertical y in [0, 31 H), columns 31 c (c = 0..W)
    if (py == 0 && cyg <= p.H && j < p.W * RW_RENDER_CELL) return black;
    if (px == 0 && cx <= p.W && gy < p.H * RW_RENDER_CELL) return black;
    if (cx >= p.W || cyg >= p.H) return white;
    const int y = p.H - 1 - cyg;  // "pyglet rendering is reversed" (186, 213, 273)
    uint32_t code = (uint32_t)(y * p.W + cx);
    // heading line from the cell centre (16, 16) to (16, 16) + 10 * heading: 10 pixels along the axis, on pixel 15 across it
    if (px == 15 && py >= 16 && py < 26) code |= 1u << 16;  // UP: +y in GL
    if (px == 15 && py >= 6 && py < 16) code |= 1u << 17;   // DOWN
    if (py == 15 && px >= 6 && px < 16) code |= 1u << 18;   // LEFT
    if (py == 15 && px >= 16 && px < 26) code |= 1u << 19;  // RIGHT
    // regular hexagon, circumradius 10, vertices at angles 0, 60, ...: |dy| <= 10 sin60 and |dy| <= sqrt(3) (10 - |dx|)
    const float dx = (float)px + 0.5f - 16.0f, dy = (float)py + 0.5f - 16.0f;
    const float ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
    if (ay <= 8.660254f && ay <= 1.7320508f * (10.0f - ax)) code |= 1u << 20;
    if (px >= 3 && px <= 28 && py >= 3 && py <= 28) code |= 1u << 21;  // shelf quad: [31 c + 3, 31 c + 29) -> pixels 3..28
    if (goal_flags[code & 0xffffu] & 2) code |= 1u << 22;              // goal quad fills the cell interior, pixels 1..30
    return code;
}

RW_HD uint32_t rw_render_shade(const RwLane &l, uint32_t code) {
    if (code >> 31) return code & 0xffffffu;
    const uint32_t gb = l.g[code & 0xffffu];
    if (gb & RW_G_AGENT) {  // agents are drawn last: heading line over hexagon over everything else
        const uint32_t aw = rw_agent(l, (int)(gb & RW_G_AGENT) - 1);
        if ((code >> (16 + rw_adir(aw))) & 1u) return rw_rgb(0, 0, 0);
        if ((code >> 20) & 1u) return rw_acarry(aw) ? rw_rgb(255, 0, 0) : rw_rgb(255, 140, 0);
    }
    if ((gb & RW_G_SHELF) && ((code >> 21) & 1u)) return (gb & RW_G_REQ) ? rw_rgb(0, 128, 128) : rw_rgb(72, 61, 139);
    if ((code >> 22) & 1u) return rw_rgb(60, 60, 60);
    return rw_rgb(255, 255, 255);
}

// colour (r | g<<8 | b<<16) of frame pixel (row i from the top, column j)
RW_HD uint32_t rw_render_pixel(const RwParams &p, const RwLane &l, const uint8_t *goal_flags, int i, int j) {
    return rw_render_shade(l, rw_render_classify(p, goal_flags, i, j));
}
