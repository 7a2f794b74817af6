// pl_dispatch on the CUDA backend: lowers a recorded pl_shader to kernel launches.
// Replaces src/dispatch.c (pl_dispatch_finish :1199-1368, translate_compute_shader :1079-1144,
// pl_dispatch_reset_frame :1613-1622): no GLSL generation, no pass cache — the "pass" is one
// of the hand-written kernels in k_*.cu parameterised by the shader's op list.
#include "shader_priv.h"

pl_dispatch pl_dispatch_create(pl_log log, pl_gpu gpu)
{
    if (!gpu) {
        plb_msg(log, PL_LOG_ERR, "pl_dispatch_create: no gpu");
        return NULL;
    }
    struct pl_dispatch_t *dp = calloc(1, sizeof(*dp));
    dp->log = log;
    dp->gpu = gpu;
    return dp;
}

void pl_dispatch_destroy(pl_dispatch *pdp)
{
    if (!pdp || !*pdp)
        return;
    struct pl_dispatch_t *dp = *pdp;
    for (int i = 0; i < dp->num_passes; i++) {
        for (int t = 0; t < PLB_PASS_TIMERS; t++)
            pl_timer_destroy(dp->gpu, &dp->passes[i]->timers[t]);
        free(dp->passes[i]);
    }
    free(dp);
    *pdp = NULL;
}

pl_shader_info pl_shader_info_ref(pl_shader_info info) { return info; }
void pl_shader_info_deref(pl_shader_info *info) { if (info) *info = NULL; }

void pl_dispatch_callback(pl_dispatch dp, void *priv,
                          void (*cb)(void *priv, const struct pl_dispatch_info *))
{
    dp->info_callback = cb;
    dp->info_priv = priv;
}

// ---- pass statistics (dispatch.c:1085-1197 `pass_info` / timer handling) ------------------
static uint64_t fnv1a(uint64_t h, const void *data, size_t n)
{
    const unsigned char *p = data;
    for (size_t i = 0; i < n; i++)
        h = (h ^ p[i]) * 0x100000001b3ull;
    return h;
}

static struct plb_pass *find_pass(pl_dispatch dp, pl_shader sh, int w, int h)
{
    uint64_t sig = 0xcbf29ce484222325ull;
    const int key[6] = { (int) sh->src, w, h, sh->px.prog.n, sh->px.nplanes, sh->is_compute };
    sig = fnv1a(sig, key, sizeof(key));
    for (int i = 0; i < sh->num_steps; i++)
        sig = fnv1a(sig, sh->steps[i], strlen(sh->steps[i]));
    for (int i = 0; i < dp->num_passes; i++) {
        if (dp->passes[i]->signature == sig)
            return dp->passes[i];
    }
    if (dp->num_passes == (int) (sizeof(dp->passes) / sizeof(dp->passes[0])))
        return NULL;
    struct plb_pass *p = calloc(1, sizeof(*p));
    if (!p)
        return NULL;
    p->signature = sig;
    // pretty description with tallies, like sh_describe's "a, b x2" (shaders.c:185-215)
    size_t len = 0;
    for (int i = 0; i < sh->num_steps; i++) {
        bool seen = false;
        for (int j = 0; j < i; j++)
            seen |= !strcmp(sh->steps[i], sh->steps[j]);
        if (seen)
            continue;
        int count = 0;
        for (int j = i; j < sh->num_steps; j++)
            count += !strcmp(sh->steps[i], sh->steps[j]);
        int n = count > 1
            ? snprintf(p->description + len, sizeof(p->description) - len, "%s%s x%d",
                       len ? ", " : "", sh->steps[i], count)
            : snprintf(p->description + len, sizeof(p->description) - len, "%s%s",
                       len ? ", " : "", sh->steps[i]);
        if (n < 0 || (size_t) n >= sizeof(p->description) - len)
            break;
        len += n;
        p->steps[p->info.num_steps++] = sh->steps[i];
    }
    // output geometry, so that tools can tell passes with the same steps apart
    snprintf(p->description + len, sizeof(p->description) - len, "%s[%dx%d]", len ? " " : "", w, h);
    p->info.params = sh->params;
    p->info.steps = p->steps;
    p->info.description = p->description;
    p->stats.shader = &p->info;
    p->stats.signature = sig;
    dp->passes[dp->num_passes++] = p;
    return p;
}

static void pass_collect(pl_dispatch dp, struct plb_pass *p)
{
    while (p->timer_tail != p->timer_head) {
        pl_timer t = p->timers[p->timer_tail % PLB_PASS_TIMERS];
        uint64_t ns = pl_timer_query(dp->gpu, t);
        if (!ns)
            break; // not finished yet (or sub-nanosecond, which cannot happen)
        p->timer_tail++;
        struct pl_dispatch_info *st = &p->stats;
        st->samples[p->sample_idx++ % 256] = ns;
        if (st->num_samples < 256)
            st->num_samples++;
        st->last = ns;
        uint64_t sum = 0, peak = 0;
        for (int i = 0; i < st->num_samples; i++) {
            sum += st->samples[i];
            peak = st->samples[i] > peak ? st->samples[i] : peak;
        }
        st->peak = peak;
        st->average = sum / st->num_samples;
    }
}

static pl_timer pass_timer(pl_dispatch dp, struct plb_pass *p)
{
    pass_collect(dp, p);
    if (p->timer_head - p->timer_tail >= PLB_PASS_TIMERS) {
        // every timer is in flight (the host runs far ahead of the GPU): wait for the oldest
        // one, so that run i of a pass always produces sample i
        plb_timer_wait(dp->gpu, p->timers[p->timer_tail % PLB_PASS_TIMERS]);
        pass_collect(dp, p);
        if (p->timer_head - p->timer_tail >= PLB_PASS_TIMERS)
            return NULL;
    }
    pl_timer *slot = &p->timers[p->timer_head % PLB_PASS_TIMERS];
    if (!*slot)
        *slot = pl_timer_create(dp->gpu);
    if (*slot)
        p->timer_head++;
    return *slot;
}

void pl_dispatch_reset_frame(pl_dispatch dp) { dp->current_index++; }
void pl_dispatch_set_frame_index(pl_dispatch dp, uint8_t index) { dp->current_index = index; }

pl_shader pl_dispatch_begin(pl_dispatch dp)
{
    pl_shader sh = pl_shader_alloc(dp->log, pl_shader_params(
        .gpu = dp->gpu,
        .index = dp->current_index,
    ));
    sh->dp = dp;
    return sh;
}

void pl_dispatch_abort(pl_dispatch dp, pl_shader *psh)
{
    (void) dp;
    pl_shader_free(psh);
}

static bool prog_has_peak(const struct plb_color_prog *prog)
{
    for (int i = 0; i < prog->n; i++) {
        if (prog->ops[i].kind == PLB_OP_PEAK)
            return true;
    }
    return false;
}

// Run the shader over `rc` of `target` (target == NULL: statistics-only dispatch of w x h)
bool plb_dispatch_run(pl_dispatch dp, pl_shader sh, pl_tex target, pl_rect2d rc, int w, int h,
                      pl_timer timer)
{
    pl_gpu gpu = dp->gpu;
    if (sh->failed) {
        PLB_ERR(dp->log, "Trying to dispatch a failed shader.");
        return false;
    }
    if (sh->src == PLB_SRC_NONE) {
        PLB_ERR(dp->log, "Trying to dispatch a shader without a source stage.");
        return false;
    }
    struct plb_pixel_launch *px = &sh->px;
    if (target) {
        if (!target->params.renderable && !target->params.storable) {
            PLB_ERR(dp->log, "pl_dispatch_finish: target texture is neither renderable nor "
                    "storable");
            return false;
        }
        if (!rc.x0 && !rc.x1) rc.x1 = target->params.w;
        if (!rc.y0 && !rc.y1) rc.y1 = target->params.h;
        if (rc.x1 < rc.x0 || rc.y1 < rc.y0) {
            PLB_ERR(dp->log, "pl_dispatch_finish: flipped target rects are not supported on "
                    "the CUDA backend");
            return false;
        }
        if (rc.x0 < 0 || rc.y0 < 0 || rc.x1 > target->params.w || rc.y1 > target->params.h) {
            PLB_ERR(dp->log, "pl_dispatch_finish: rect out of target bounds");
            return false;
        }
        w = rc.x1 - rc.x0;
        h = rc.y1 - rc.y0;
        px->dst = plb_tex_priv(target)->d;
    } else {
        memset(&px->dst, 0, sizeof(px->dst));
        rc = (pl_rect2d) { 0, 0, w, h };
    }
    if ((sh->out_w && sh->out_w != w) || (sh->out_h && sh->out_h != h)) {
        PLB_ERR(dp->log, "Trying to dispatch a shader with incompatible size %dx%d to a "
                "region of size %dx%d!", sh->out_w, sh->out_h, w, h);
        return false;
    }
    px->dst_rect = rc;
    px->out_w = w;
    px->out_h = h;
    px->frame_index = sh->params.index;
    // work-group geometry as the reference's merged compute shader would have it:
    // rigid 32x32 from the polar sampler, else the flexible 16x16 of pl_shader_detect_peak
    px->wg_w = sh->is_compute ? sh->wg_w : 16;
    px->wg_h = sh->is_compute ? sh->wg_h : 16;
    if (prog_has_peak(&px->prog) && (32 % px->wg_w || 32 % px->wg_h)) {
        // our tiles hold whole work groups of 8/16/32; anything else falls back to 16x16,
        // which only changes the statistics within the reference's own tolerance
        px->wg_w = px->wg_h = 16;
    }

    // internal timing only when the caller did not claim the interval (dispatch.h:111-117)
    struct plb_pass *pass = NULL;
    if (!timer && dp->info_callback && !plb_gpu_no_timers(gpu)) {
        pass = find_pass(dp, sh, w, h);
        if (pass)
            timer = pass_timer(dp, pass);
    } else if (dp->info_callback) {
        pass = find_pass(dp, sh, w, h);
    }

    // order this launch after pending host<->device transfers of its textures
    pl_tex texs[13];
    int ntexs = 0;
    for (int i = 0; i < sh->num_texs; i++)
        texs[ntexs++] = sh->texs[i];
    if (target)
        texs[ntexs++] = target;
    plb_gpu_before_use(gpu, texs, ntexs);

    extern int plb_gpu_enter(pl_gpu gpu);
    extern void plb_gpu_leave(int prev);
    const int prev_dev = plb_gpu_enter(gpu); // launches go to the gpu's device, whatever is current
    plb_timer_begin(gpu, timer);
    bool ok = false;
    switch (sh->src) {
    case PLB_SRC_PLANES:
        ok = plb_launch_pixel(gpu, px);
        break;
    case PLB_SRC_POLAR:
        sh->polar.tail = *px;
        ok = plb_launch_polar(gpu, &sh->polar);
        break;
    case PLB_SRC_ORTHO:
        sh->ortho.tail = *px;
        ok = plb_launch_ortho(gpu, &sh->ortho);
        break;
    case PLB_SRC_ERRDIFF:
        ok = plb_launch_errdiff(gpu, &sh->errdiff);
        break;
    case PLB_SRC_MIX:
        sh->mix.tail = *px;
        ok = plb_launch_mix(gpu, &sh->mix);
        break;
    case PLB_SRC_DEINT:
        sh->deint.tail = *px;
        ok = plb_launch_deint(gpu, &sh->deint);
        break;
    default:
        break;
    }
    plb_timer_end(gpu, timer);
    plb_gpu_after_use(gpu, texs, ntexs);
    plb_gpu_leave(prev_dev);
    if (ok && pass && dp->info_callback) {
        pass_collect(dp, pass);
        dp->info_callback(dp->info_priv, &pass->stats);
    }
    return ok;
}

bool pl_dispatch_finish(pl_dispatch dp, const struct pl_dispatch_params *params)
{
    pl_shader sh = params && params->shader ? *params->shader : NULL;
    bool ok = false;
    if (!sh) {
        PLB_ERR(dp->log, "pl_dispatch_finish: no shader");
        return false;
    }
    if (!params->target) {
        PLB_ERR(dp->log, "pl_dispatch_finish: no target texture");
        goto done;
    }
    if (params->blend_params) {
        PLB_ERR(dp->log, "pl_dispatch_finish: blending is outside the CUDA hot path");
        goto done;
    }
    if (sh->src == PLB_SRC_ERRDIFF) {
        PLB_ERR(dp->log, "Trying to dispatch a compute-only shader with pl_dispatch_finish; "
                "use pl_dispatch_compute");
        goto done;
    }
    ok = plb_dispatch_run(dp, sh, params->target, params->rect, 0, 0, params->timer);
done:
    // the dispatch consumes the shader either way (dispatch.h:84-90)
    pl_shader_free(params->shader);
    return ok;
}

bool pl_dispatch_compute(pl_dispatch dp, const struct pl_dispatch_compute_params *params)
{
    pl_shader sh = params && params->shader ? *params->shader : NULL;
    bool ok = false;
    if (!sh) {
        PLB_ERR(dp->log, "pl_dispatch_compute: no shader");
        return false;
    }
    if (sh->src == PLB_SRC_ERRDIFF) {
        // the error diffusion shader carries its own textures (dithering.c:326-406)
        ok = plb_dispatch_run(dp, sh, NULL, (pl_rect2d) {0}, sh->errdiff.w, sh->errdiff.h,
                              params->timer);
    } else if (params->width > 0 && params->height > 0) {
        // output-less compute pass (e.g. a stand-alone peak detection run)
        ok = plb_dispatch_run(dp, sh, NULL, (pl_rect2d) {0}, params->width, params->height,
                              params->timer);
    } else {
        PLB_ERR(dp->log, "pl_dispatch_compute: width/height required on the CUDA backend");
    }
    pl_shader_free(params->shader);
    return ok;
}
