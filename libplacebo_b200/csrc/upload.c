// Frame ingest helpers (libplacebo/utils/upload.h, src/utils/upload.c): describe a plane of host
// or buffer pixels bit by bit (`pl_plane_data`), find the texture format whose host layout is
// exactly that (`pl_plane_find_fmt`), and (re)create + fill the plane texture
// (`pl_upload_plane`, `pl_recreate_plane`). Pure host logic on top of pl_tex_recreate /
// pl_tex_upload; the copies themselves run on the copy engines (gpu.cu).
#include <assert.h>
#include <stdlib.h>

#include "internal.h"

enum { NCOMP = 4 };

// One colour channel of a packed pixel: bits [lsb, lsb + width)
struct chan {
    int index;  // 0..3 = R, G, B, A
    int width;
    int lsb;
};

static int chan_order(const void *pa, const void *pb)
{
    const struct chan *a = pa, *b = pb;
    // present channels first, then by position in the pixel
    if (!a->width != !b->width)
        return a->width ? -1 : 1;
    return (a->lsb > b->lsb) - (a->lsb < b->lsb);
}

// upload.c:46-75: (size, shift) per channel -> the sequence "pad, size, pad, size, ..." in
// memory order plus the channel each slot carries
void pl_plane_data_from_comps(struct pl_plane_data *data, int size[4], int shift[4])
{
    struct chan ch[NCOMP];
    for (int i = 0; i < NCOMP; i++)
        ch[i] = (struct chan) { .index = i, .width = size[i], .lsb = shift[i] };
    qsort(ch, NCOMP, sizeof(ch[0]), chan_order);

    int cursor = 0; // first bit not yet described
    for (int i = 0; i < NCOMP; i++) {
        if (!ch[i].width) {
            data->component_size[i] = data->component_pad[i] = data->component_map[i] = 0;
            continue;
        }
        assert(ch[i].lsb >= cursor);
        data->component_size[i] = ch[i].width;
        data->component_pad[i] = ch[i].lsb - cursor;
        data->component_map[i] = ch[i].index;
        cursor = ch[i].lsb + ch[i].width;
    }
}

// upload.c:77-93: contiguous bit masks -> (size, shift)
void pl_plane_data_from_mask(struct pl_plane_data *data, uint64_t mask[4])
{
    int size[NCOMP], shift[NCOMP];
    for (int i = 0; i < NCOMP; i++) {
        size[i] = __builtin_popcountll(mask[i]);
        shift[i] = mask[i] ? __builtin_ctzll(mask[i]) : 0;
        assert(size[i] == 64 ? mask[i] == ~0ull : mask[i] == (((1ull << size[i]) - 1) << shift[i]));
    }
    pl_plane_data_from_comps(data, size, shift);
}

static int round_down8(int x) { return x & ~7; } // x >= 0 here
static int round_up8(int x) { return (x + 7) & ~7; }

// upload.c:95-158: widen every component to whole bytes by absorbing the padding around it;
// possible only if all components end up with the same depth / shift / sample size, which is
// then the pl_bit_encoding of the plane (e.g. P010: 10 bits in the top of 16 -> shift 6).
bool pl_plane_data_align(struct pl_plane_data *data, struct pl_bit_encoding *out_bits)
{
    struct pl_plane_data d = *data;
    struct pl_bit_encoding enc = {0};
    int cursor = 0;
    for (int i = 0; i < NCOMP && d.component_size[i]; i++) {
        const bool first = cursor == 0;
        const int depth = d.component_size[i];
        const int begin = cursor + d.component_pad[i];
        const int end = begin + depth;
        // low side: down to the byte boundary, but never beyond this component's own padding
        int grow_lo = begin - round_down8(begin);
        if (grow_lo > d.component_pad[i])
            grow_lo = d.component_pad[i];
        // high side: up to the byte boundary; the last component may take it freely, the
        // others only from the padding of their successor
        int grow_hi = round_up8(end) - end;
        const bool last = i + 1 == NCOMP || !d.component_size[i + 1];
        if (!last) {
            if (grow_hi > d.component_pad[i + 1])
                grow_hi = d.component_pad[i + 1];
            d.component_pad[i + 1] -= grow_hi;
        }
        d.component_pad[i] -= grow_lo;
        d.component_size[i] = depth + grow_lo + grow_hi;

        if (first) {
            enc.color_depth = depth;
            enc.bit_shift = grow_lo;
            enc.sample_depth = d.component_size[i];
        } else if (enc.color_depth != depth || enc.bit_shift != grow_lo ||
                   enc.sample_depth != d.component_size[i]) {
            goto fail;
        }
        cursor += d.component_pad[i] + d.component_size[i];
    }
    if (d.pixel_stride && cursor > (int) d.pixel_stride * 8)
        goto fail;
    *data = d;
    if (out_bits)
        *out_bits = enc;
    return true;

fail:
    if (out_bits)
        *out_bits = (struct pl_bit_encoding) {0};
    return false;
}

// Does `fmt` store exactly the pad/size sequence of `data`? Records the channel carried by each
// texture component as it goes (a padding slot consumes a component of its own). Like the
// reference (upload.c:184-196) the entries written for a format that is rejected later are NOT
// rolled back, so `map` is only meaningful up to the components of the format returned.
static bool layout_matches(pl_fmt fmt, const struct pl_plane_data *data, int used, int map[4])
{
    int slot = 0;
    for (int i = 0; i < used; i++) {
        const int pad = data->component_pad[i], size = data->component_size[i];
        if (pad) {
            if (slot >= NCOMP || fmt->host_bits[slot] != pad)
                return false;
            slot++;
        }
        if (slot >= NCOMP || (size && fmt->host_bits[slot] != size))
            return false;
        map[slot++] = data->component_map[i];
    }
    return true;
}

// upload.c:160-214: the first sampleable format (in list order) of the right type and pixel
// size whose host layout matches
pl_fmt pl_plane_find_fmt(pl_gpu gpu, int out_map[4], const struct pl_plane_data *data)
{
    int scratch[4];
    int *map = out_map ? out_map : scratch;
    if (data->swapped && !gpu->limits.max_ssbo_size)
        return NULL; // byte swapping runs as a compute pass over a storage buffer
    int used = 0;
    for (int i = 0; i < NCOMP; i++) {
        map[i] = -1;
        if (data->component_size[i])
            used = i + 1;
    }
    for (int n = 0; n < gpu->num_formats; n++) {
        pl_fmt fmt = gpu->formats[n];
        if (fmt->opaque || fmt->num_components < used || fmt->type != data->type ||
            fmt->texel_size != data->pixel_stride || !(fmt->caps & PL_FMT_CAP_SAMPLEABLE))
            continue;
        if (!layout_matches(fmt, data, used, map))
            continue;
        if (data->row_stride % fmt->texel_align) {
            PLB_WARN(gpu->log, "pl_plane_find_fmt: rejecting '%s': row stride %zu is not a "
                     "multiple of the texel alignment %zu (API usage bug?)", fmt->name,
                     data->row_stride, fmt->texel_align);
            continue;
        }
        return fmt;
    }
    return NULL;
}

static void describe_plane(struct pl_plane *out, pl_tex tex, const int map[4])
{
    if (!out)
        return;
    out->texture = tex;
    out->components = 0;
    for (int i = 0; i < NCOMP; i++) {
        out->component_mapping[i] = map[i];
        if (map[i] >= 0)
            out->components = i + 1;
    }
}

// upload.c:216-319
bool pl_upload_plane(pl_gpu gpu, struct pl_plane *out_plane, pl_tex *tex,
                     const struct pl_plane_data *data)
{
    if (!data || !tex || (!data->buf == !data->pixels)) {
        PLB_ERR(gpu->log, "pl_upload_plane: exactly one of `pixels` and `buf` must be set");
        return false;
    }
    int map[4];
    pl_fmt fmt = pl_plane_find_fmt(gpu, map, data);
    if (!fmt) {
        PLB_ERR(gpu->log, "Failed picking any compatible texture format for a plane!");
        return false;
    }
    if (!pl_tex_recreate(gpu, tex, pl_tex_params(
            .w = data->width,
            .h = data->height,
            .format = fmt,
            .sampleable = true,
            .host_writable = true,
            .blit_src = !!(fmt->caps & PL_FMT_CAP_BLITTABLE),
        ))) {
        PLB_ERR(gpu->log, "Failed initializing plane texture!");
        return false;
    }
    describe_plane(out_plane, *tex, map);

    struct pl_tex_transfer_params xfer = {
        .tex = *tex,
        .rc = { .x1 = data->width, .y1 = data->height, .z1 = 1 },
        .row_pitch = data->row_stride ? data->row_stride : data->width * fmt->texel_size,
        .ptr = (void *) data->pixels,
        .buf = data->buf,
        .buf_offset = data->buf_offset,
        .callback = data->callback,
        .priv = data->priv,
    };
    if (!data->swapped)
        return pl_tex_upload(gpu, &xfer);

    // Foreign-endian samples: through a storage buffer that a byte-swapping kernel rewrites in
    // place (pl_buf_copy_swap), then a buffer -> texture copy.
    // minimal size of the transfer (pl_tex_transfer_size, gpu/utils.c:472-481), in whole words
    const size_t exact = (size_t) (data->height - 1) * xfer.row_pitch + (size_t) data->width * fmt->texel_size;
    const size_t bytes = (exact + 3) & ~(size_t) 3;
    pl_buf tmp = pl_buf_create(gpu, pl_buf_params( .size = bytes, .storable = true ));
    if (!tmp) {
        PLB_ERR(gpu->log, "Failed creating endian swapping buffer!");
        return false;
    }
    pl_buf src = tmp;
    size_t src_offset = 0;
    if (xfer.ptr) {
        pl_buf_write(gpu, tmp, 0, xfer.ptr, exact);
    } else if (xfer.buf->params.storable && xfer.buf_offset % 4 == 0 &&
               xfer.buf_offset + bytes <= xfer.buf->params.size) {
        src = xfer.buf; // swap straight out of the caller's buffer
        src_offset = xfer.buf_offset;
    } else {
        const size_t avail = xfer.buf->params.size - xfer.buf_offset;
        pl_buf_copy(gpu, tmp, 0, xfer.buf, xfer.buf_offset, avail < bytes ? avail : bytes);
    }
    bool ok = plb_buf_copy_swap(gpu, tmp, 0, src, src_offset, bytes,
                                (int) (fmt->texel_size / fmt->num_components));
    if (!ok)
        PLB_ERR(gpu->log, "Failed swapping endianness!");
    xfer.ptr = NULL;
    xfer.buf = tmp;
    xfer.buf_offset = 0;
    ok = ok && pl_tex_upload(gpu, &xfer);
    pl_buf_destroy(gpu, &tmp);
    return ok;
}

// upload.c:321-360: the same format search, but the texture is created as a render target
bool pl_recreate_plane(pl_gpu gpu, struct pl_plane *out_plane, pl_tex *tex,
                       const struct pl_plane_data *data)
{
    if (data->swapped) {
        PLB_ERR(gpu->log, "Cannot call pl_recreate_plane on non-native endian plane data, this "
                "is only supported for `pl_upload_plane`!");
        return false;
    }
    int map[4];
    pl_fmt fmt = pl_plane_find_fmt(gpu, map, data);
    if (!fmt) {
        PLB_ERR(gpu->log, "Failed picking any compatible texture format for a plane!");
        return false;
    }
    if (!pl_tex_recreate(gpu, tex, pl_tex_params(
            .w = data->width,
            .h = data->height,
            .format = fmt,
            .renderable = true,
            .host_readable = !!(fmt->caps & PL_FMT_CAP_HOST_READABLE),
            .blit_dst = !!(fmt->caps & PL_FMT_CAP_BLITTABLE),
            .storable = !!(fmt->caps & PL_FMT_CAP_STORABLE),
        ))) {
        PLB_ERR(gpu->log, "Failed initializing plane texture!");
        return false;
    }
    describe_plane(out_plane, *tex, map);
    return true;
}
