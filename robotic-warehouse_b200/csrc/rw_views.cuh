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

// value of layer `layer` at row y, column x (0 <= y < H, 0 <= x < W).  goal_flags[cell] & 2 marks goal cells.
RW_HD float rw_gimg_value(const RwParams &p, const RwLane &l, const uint8_t *goal_flags, int layer, int y, int x) {
    const int cell = y * p.W + x;
    const uint32_t gb = l.g[cell];
    switch (layer) {
        case 0: return (gb & RW_G_SHELF) ? 1.0f : 0.0f;   // SHELVES 988-991
        case 1: return (gb & RW_G_REQ) ? 1.0f : 0.0f;     // REQUESTS 992-995
        case 2: return (gb & RW_G_AGENT) ? 1.0f : 0.0f;   // AGENTS 996-999
        case 5: return (goal_flags[cell] & 2) ? 1.0f : 0.0f;  // GOALS 1010-1013 (the swapped loop names cancel: layer[y][x])
        case 6: return (gb & RW_G_AGENT) ? 0.0f : 1.0f;   // ACCESSIBLE 1014-1017
        case 3:
        case 4: {  // layer[ag.x, ag.y] (sic): element [y][x] belongs to the agent standing at column y, row x
            if (x >= p.H || y >= p.W) return 0.0f;
            uint32_t id = l.g[x * p.W + y] & RW_G_AGENT;
            if (!id) return 0.0f;
            uint32_t aw = rw_agent(l, (int)id - 1);
            if (layer == 3) return (float)(rw_adir(aw) + 1);  // AGENT_DIRECTION 1000-1004
            return rw_acarry(aw) ? 1.0f : 0.0f;               // AGENT_LOAD 1005-1009
        }
    }
    return 0.0f;
}

RW_HD float rw_gimg_element(const RwParams &p, const RwLane &l, const uint8_t *goal_flags, const RwViewSpec &s, int idx) {
    const int plane = s.Hp * s.Wp;
    const int cp = idx / plane, rem = idx - cp * plane, yp = rem / s.Wp, xp = rem - yp * s.Wp;
    const int c = cp - s.c0, y = yp - s.y0, x = xp - s.x0;
    if (c < 0 || c >= s.n_layers || y < 0 || y >= p.H || x < 0 || x >= p.W) return 0.0f;  // np.pad constant 0 (1033-1038)
    return rw_gimg_value(p, l, goal_flags, s.layers[c], y, x);
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

// colour (r | g<<8 | b<<16) of frame pixel (row i from the top, column j)
RW_HD uint32_t rw_render_pixel(const RwParams &p, const RwLane &l, const uint8_t *goal_flags, int i, int j) {
    const uint32_t white = rw_rgb(255, 255, 255), black = rw_rgb(0, 0, 0);
    const int gy = rw_render_height(p) - 1 - i;  // OpenGL row (origin bottom-left)
    const int cx = j / RW_RENDER_CELL, px = j - cx * RW_RENDER_CELL;  // px == 0: vertical grid line
    const int cyg = gy / RW_RENDER_CELL, py = gy - cyg * RW_RENDER_CELL;  // py == 0: horizontal grid line
    const bool in_x = cx < p.W, in_y = cyg < p.H;
    // grid lines: horizontal x in [0, 31 W), rows 31 r (r = 0..H); vertical y in [0, 31 H), columns 31 c (c = 0..W)
    if (py == 0 && cyg <= p.H && j < p.W * RW_RENDER_CELL) return black;
    if (px == 0 && cx <= p.W && gy < p.H * RW_RENDER_CELL) return black;
    if (!in_x || !in_y) return white;
    const int y = p.H - 1 - cyg;  // "pyglet rendering is reversed" (186, 213, 273)
    const int cell = y * p.W + cx;
    const uint32_t gb = l.g[cell];
    // agent: heading line, then hexagon, both relative to the centre (16, 16) of the 31-pixel cell, pixel centre (px+.5, py+.5)
    if (gb & RW_G_AGENT) {
        const uint32_t aw = rw_agent(l, (int)(gb & RW_G_AGENT) - 1);
        const int d = rw_adir(aw);
        // line from (16,16) to (16,16) + 10 * heading: covers 10 pixels along the axis, on pixel 15 across it
        const bool on_line = d == 0   ? (px == 15 && py >= 16 && py < 26)    // UP: +y in GL
                             : d == 1 ? (px == 15 && py >= 6 && py < 16)     // DOWN
                             : d == 2 ? (py == 15 && px >= 6 && px < 16)     // LEFT
                                      : (py == 15 && px >= 16 && px < 26);   // RIGHT
        if (on_line) return black;
        // regular hexagon, circumradius 10, vertices at angles 0, 60, ...: |dy| <= 10 sin60 and |dy| <= sqrt(3) (10 - |dx|)
        const float dx = (float)px + 0.5f - 16.0f, dy = (float)py + 0.5f - 16.0f;
        const float ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
        if (ay <= 8.660254f && ay <= 1.7320508f * (10.0f - ax)) return rw_acarry(aw) ? rw_rgb(255, 0, 0) : rw_rgb(255, 140, 0);
    }
    // shelf quad: x, y in [31 c + 3, 31 c + 29) continuous -> pixels 3..28
    if ((gb & RW_G_SHELF) && px >= 3 && px <= 28 && py >= 3 && py <= 28)
        return (gb & RW_G_REQ) ? rw_rgb(0, 128, 128) : rw_rgb(72, 61, 139);
    if (goal_flags[cell] & 2) return rw_rgb(60, 60, 60);  // goal quad fills the cell interior, pixels 1..30
    return white;
}
