/*
 * oracle/age_oracle.c -- TEST INFRASTRUCTURE ONLY (see age_oracle.h).
 *
 * Plain-C restatement of one single-mode AGE aligner as implemented by the reference
 * (all citations are /root/reference/src/AsmvarAGE/src/AGEaligner.cpp unless noted).
 * It keeps the reference's *semantics* -- full (len2+2) x (len1+2) u16 score matrices with
 * zero borders, one 2-bit field per trace plane -- but is written independently: four
 * separate planes instead of one packed byte, explicit [r][c] indexing instead of
 * running pointers, no OpenMP.
 *
 * Conventions: seq1 = first sequence (target window), columns c = 1..len1;
 *              seq2 = second sequence (contig),       rows    r = 1..len2.
 */
#include "age_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

enum { T_STOP = 0, T_DIAG = 1, T_HORI = 2, T_VERT = 3 };   /* AGEaligner.hh:41-44 */

char orc_complement(char c)
{   /* Sequence.hh:138-150: only acgtu/ACGTU change; u->a, U->A */
    switch (c) {
    case 'a': return 't'; case 'c': return 'g'; case 't': return 'a'; case 'g': return 'c';
    case 'A': return 'T'; case 'C': return 'G'; case 'T': return 'A'; case 'G': return 'C';
    case 'u': return 'a'; case 'U': return 'A';
    default:  return c;
    }
}

void orc_revcom(const char *in, int n, char *out)
{   /* Sequence.hh:195-197 */
    for (int i = 0; i < n; i++) out[i] = orc_complement(in[n - 1 - i]);
}

static int same_nuc(char a, char b)
{   /* Sequence.hh:182-188 */
    int aa = toupper((unsigned char)a), bb = toupper((unsigned char)b);
    int va = (aa == 'A' || aa == 'C' || aa == 'T' || aa == 'G' || aa == 'U');
    int vb = (bb == 'A' || bb == 'C' || bb == 'T' || bb == 'G' || bb == 'U');
    return va && vb && aa == bb;
}

int orc_subst_score(char a, char b, int match, int mismatch)
{   /* Scorer.hh:24-39: characters that complement() leaves unchanged score 0 */
    if (orc_complement(a) == a) return 0;
    if (orc_complement(b) == b) return 0;
    return same_nuc(a, b) ? match : mismatch;
}

typedef struct {
    int len1, len2, n1, n2;          /* n1 = len1+2, n2 = len2+2  (:29-30) */
    uint16_t *F, *R;                 /* score matrices, later overwritten by their maxima */
    uint8_t  *FS, *RS, *FM, *RM;     /* the four 2-bit trace fields of AGEaligner.hh:46-64 */
    int flag;
    int bp[ORC_MAX_BP][4];
    int n_bp;
} mat_t;

#define AT(m, a, r, c) ((m)->a[(size_t)(r) * (size_t)(m)->n1 + (size_t)(c)])

static int gap_cost(int tr, int go, int ge)
{   /* :1006-1013 -- the penalty for leaving a cell depends on that cell's own trace */
    return (tr == T_HORI || tr == T_VERT) ? ge : go;
}

/* K1/K2: forward and reverse local fills (:988-1028, :1034-1078) */
static void fill_scores(mat_t *m, const char *s1, const char *s1rc, const char *s2,
                        int match, int mismatch, int go, int ge)
{
    const char *xf = (m->flag & ORC_INVR) ? s1rc : s1;   /* :1000 */
    const char *xr = (m->flag & ORC_INVL) ? s1rc : s1;   /* :1050 */
    for (int r = 1; r <= m->len2; r++)
        for (int c = 1; c <= m->len1; c++) {
            int d = AT(m, F, r - 1, c - 1) + orc_subst_score(s2[r - 1], xf[c - 1], match, mismatch);
            int h = AT(m, F, r, c - 1) + gap_cost(AT(m, FS, r, c - 1), go, ge);
            int v = AT(m, F, r - 1, c) + gap_cost(AT(m, FS, r - 1, c), go, ge);
            int s = d, t = T_DIAG;
            if (h >= s) { s = h; t = T_HORI; }            /* :1015 */
            if (v >= s) { s = v; t = T_VERT; }            /* :1016 */
            if (s < 0)  { s = 0; t = T_STOP; }            /* :1017 */
            AT(m, F, r, c)  = (uint16_t)s;                /* :1019 (u16 store) */
            AT(m, FS, r, c) = (uint8_t)t;
        }
    for (int r = m->len2; r >= 1; r--)
        for (int c = m->len1; c >= 1; c--) {
            int d = AT(m, R, r + 1, c + 1) + orc_subst_score(s2[r - 1], xr[c - 1], match, mismatch);
            int h = AT(m, R, r, c + 1) + gap_cost(AT(m, RS, r, c + 1), go, ge);
            int v = AT(m, R, r + 1, c) + gap_cost(AT(m, RS, r + 1, c), go, ge);
            int s = d, t = T_DIAG;
            if (h >= s) { s = h; t = T_HORI; }
            if (v >= s) { s = v; t = T_VERT; }
            if (s < 0)  { s = 0; t = T_STOP; }
            AT(m, R, r, c)  = (uint16_t)s;
            AT(m, RS, r, c) = (uint8_t)t;
        }
}

/* K3/K4: in-place prefix / suffix maxima with their trace (:873-917, :923-970) */
static void fill_maxima(mat_t *m)
{
    for (int c = 1; c <= m->len1 + 1; c++) AT(m, FM, 0, c) = T_HORI;   /* :874-878 */
    for (int r = 1; r <= m->len2; r++)     AT(m, FM, r, 0) = T_VERT;   /* :879-882 */
    for (int r = 1; r <= m->len2; r++)
        for (int c = 1; c <= m->len1; c++) {
            int t = T_STOP;
            if (AT(m, F, r - 1, c) > AT(m, F, r, c) && r > 1) { AT(m, F, r, c) = AT(m, F, r - 1, c); t = T_VERT; }
            if (AT(m, F, r, c - 1) > AT(m, F, r, c) && c > 1) { AT(m, F, r, c) = AT(m, F, r, c - 1); t = T_HORI; }
            AT(m, FM, r, c) = (uint8_t)t;
        }
    /* Reverse border: the reference tags row len2+1 / column len1+1 with *FM* bits (:924-932),
     * so RM stays 0 there.  Those stray FM bits are never read; we simply leave RM = 0. */
    for (int r = m->len2; r >= 1; r--)
        for (int c = m->len1; c >= 1; c--) {
            int t = T_STOP;
            if (AT(m, R, r + 1, c) > AT(m, R, r, c) && r < m->len2) { AT(m, R, r, c) = AT(m, R, r + 1, c); t = T_VERT; }
            if (AT(m, R, r, c + 1) > AT(m, R, r, c) && c < m->len1) { AT(m, R, r, c) = AT(m, R, r, c + 1); t = T_HORI; }
            AT(m, RM, r, c) = (uint8_t)t;
        }
}

/* K8 traceBackMaxima (:615-642) */
static void walk_maxima(const mat_t *m, int *lg1, int *lg2, int *rg1, int *rg2)
{
    int t;
    while ((t = AT(m, FM, *lg2, *lg1)) != T_STOP) { if (t == T_VERT) (*lg2)--; else (*lg1)--; }
    while ((t = AT(m, RM, *rg2, *rg1)) != T_STOP) { if (t == T_VERT) (*rg2)++; else (*rg1)++; }
}

/* K9 addBpoints (:644-668) -- note the reference tests lg2 (not rg2) against the last row */
static int add_bp(mat_t *m, int lg1, int lg2, int rg1, int rg2)
{
    if (lg1 == 0 || lg2 == 0) lg1 = lg2 = 0;
    if (rg1 == m->n1 - 1 || lg2 == m->n2 - 1) { rg1 = m->n1 - 1; rg2 = m->n2 - 1; }
    for (int i = 0; i < m->n_bp; i++)
        if (m->bp[i][0] - lg1 == m->bp[i][2] - rg1 && m->bp[i][1] - lg2 == m->bp[i][3] - rg2) return 0;
    if (m->n_bp < ORC_MAX_BP) {
        int *b = m->bp[m->n_bp++];
        b[0] = lg1; b[1] = lg2; b[2] = rg1; b[3] = rg2;
        return 1;
    }
    return -1;
}

static int try_candidate(mat_t *m, int lg1, int lg2, int rg1, int rg2, int *n_add, int max_bps)
{   /* returns 1 when the enumeration must stop (:565-566) */
    walk_maxima(m, &lg1, &lg2, &rg1, &rg2);
    int res = add_bp(m, lg1, lg2, rg1, rg2);
    if (res > 0) (*n_add)++;
    return (*n_add >= max_bps || res < 0);
}

/* K6 findIndelBPs (:540-613) */
static int find_indel(mat_t *m, int forward, int max_bps)
{
    int max = 0, n_add = 0;
    for (int r = 0; r <= m->len2; r++)
        for (int c = 0; c <= m->len1; c++) {
            uint16_t val = (uint16_t)(AT(m, F, r, c) + AT(m, R, r + 1, c + 1));   /* u16 wrap, :548 */
            if (val > max) max = val;
        }
    if (forward) {
        for (int r = 0; r <= m->len2; r++)
            for (int c = 0; c <= m->len1; c++)
                if (AT(m, FM, r, c) == T_STOP && AT(m, F, r, c) + AT(m, R, r + 1, c + 1) == max)
                    if (try_candidate(m, c, r, c + 1, r + 1, &n_add, max_bps)) return max;
    } else {
        for (int r = m->len2; r >= 0; r--)
            for (int c = m->len1; c >= 0; c--)
                if (AT(m, RM, r + 1, c + 1) == T_STOP && AT(m, F, r, c) + AT(m, R, r + 1, c + 1) == max)
                    if (try_candidate(m, c, r, c + 1, r + 1, &n_add, max_bps)) return max;
    }
    return max;
}

/* K7 findInversionTDuplicationBPs (:472-538) */
static int find_inv_tdup(mat_t *m, int forward, int max_bps)
{
    int max = 0, n_add = 0, L1 = m->len1;
    for (int r = 0; r <= m->len2; r++) {
        uint16_t val = (uint16_t)(AT(m, F, r, L1) + AT(m, R, r + 1, 1));           /* :479 */
        if (val > max) max = val;
    }
    if (forward) {
        for (int r = 0; r <= m->len2; r++) {
            if (AT(m, F, r, L1) + AT(m, R, r + 1, 1) != max) continue;
            for (int j1 = 0; j1 <= L1; j1++) {
                if (AT(m, FM, r, j1) != T_STOP) continue;
                for (int j2 = 0; j2 <= L1; j2++)
                    if (AT(m, F, r, j1) + AT(m, R, r + 1, j2 + 1) == max)
                        if (try_candidate(m, j1, r, j2 + 1, r + 1, &n_add, max_bps)) return max;
            }
        }
    } else {
        for (int i2 = m->len2 + 1; i2 >= 1; i2--) {
            if (AT(m, F, i2 - 1, L1) + AT(m, R, i2, 1) != max) continue;
            for (int j2 = L1 + 1; j2 >= 1; j2--) {
                if (AT(m, RM, i2, j2) != T_STOP) continue;
                for (int j1 = L1 + 1; j1 >= 1; j1--)
                    if (AT(m, F, i2 - 1, j1 - 1) + AT(m, R, i2, j2) == max)
                        if (try_candidate(m, j1 - 1, i2 - 1, j2, i2, &n_add, max_bps)) return max;
            }
        }
    }
    return max;
}

/* K11 findLeftBlocks (:752-808): follow FS back from (left2,left1), merge diagonal runs */
static int left_blocks(const mat_t *m, int c, int r, orc_block **out)
{
    int cap = (m->len1 > m->len2 ? m->len1 : m->len2) + 1, n = 0;
    orc_block *b = (orc_block *)malloc(sizeof(orc_block) * (size_t)cap);
    int t, run = 0, last_c = 0, last_r = 0;
    /* walking backwards: a diagonal run ends (in walk order) at its lowest coordinate */
    while ((t = AT(m, FS, r, c)) != T_STOP) {
        if (t == T_DIAG) {
            if (run > 0 && last_c - c == 1 && last_r - r == 1) run++;
            else {
                if (run > 0) { b[n].start1 = last_c; b[n].start2 = last_r; b[n].length = run; n++; }
                run = 1;
            }
            last_c = c; last_r = r;
            c--; r--;
        } else if (t == T_HORI) c--;
        else r--;
    }
    if (run > 0) { b[n].start1 = last_c; b[n].start2 = last_r; b[n].length = run; n++; }
    for (int i = 0; i < n / 2; i++) { orc_block tmp = b[i]; b[i] = b[n - 1 - i]; b[n - 1 - i] = tmp; }
    *out = b;
    return n;
}

/* K12 findRightBlocks (:810-860) */
static int right_blocks(const mat_t *m, int c, int r, orc_block **out)
{
    int cap = (m->len1 > m->len2 ? m->len1 : m->len2) + 1, n = 0;
    orc_block *b = (orc_block *)malloc(sizeof(orc_block) * (size_t)cap);
    int t, run = 0, s1 = 0, s2 = 0;
    while ((t = AT(m, RS, r, c)) != T_STOP) {
        if (t == T_DIAG) {
            if (c <= m->len1 && r <= m->len2) {           /* :820 */
                if (run == 0) { s1 = c; s2 = r; run = 1; } else run++;
            }
            c++; r++;
        } else {
            if (t == T_HORI) c++; else r++;
            if (run > 0) { b[n].start1 = s1; b[n].start2 = s2; b[n].length = run; n++; run = 0; }
        }
    }
    if (run > 0) { b[n].start1 = s1; b[n].start2 = s2; b[n].length = run; n++; }
    *out = b;
    return n;
}

static int alloc_mat(mat_t *m, int len1, int len2, int flag)
{
    memset(m, 0, sizeof(*m));
    m->len1 = len1; m->len2 = len2; m->n1 = len1 + 2; m->n2 = len2 + 2; m->flag = flag;
    size_t sz = (size_t)m->n1 * (size_t)m->n2;
    m->F = (uint16_t *)calloc(sz, 2); m->R = (uint16_t *)calloc(sz, 2);
    m->FS = (uint8_t *)calloc(sz, 1); m->RS = (uint8_t *)calloc(sz, 1);
    m->FM = (uint8_t *)calloc(sz, 1); m->RM = (uint8_t *)calloc(sz, 1);
    return (m->F && m->R && m->FS && m->RS && m->FM && m->RM) ? 0 : -1;
}

static void free_mat(mat_t *m)
{
    free(m->F); free(m->R); free(m->FS); free(m->RS); free(m->FM); free(m->RM);
}

static int check_flag(int flag)
{
    return flag == ORC_INDEL || flag == ORC_TDUP || flag == ORC_INVL || flag == ORC_INVR;
}

int orc_align(const char *seq1, int len1, const char *seq2, int len2,
              int flag, int match, int mismatch, int go, int ge, orc_result *out)
{
    memset(out, 0, sizeof(*out));
    out->alt_index = -1;
    if (len1 <= 0 || len2 <= 0) return -1;                /* :68 */
    if (!check_flag(flag)) return -1;                     /* :69 (single-mode restatement) */
    mat_t m;
    if (alloc_mat(&m, len1, len2, flag)) return -2;
    char *rc = (char *)malloc((size_t)len1 + 1);
    orc_revcom(seq1, len1, rc); rc[len1] = 0;             /* :27 */

    fill_scores(&m, seq1, rc, seq2, (short)match, (short)mismatch, (short)go, (short)ge);
    fill_maxima(&m);
    int max;
    if (flag == ORC_INDEL) { max = find_indel(&m, 1, ORC_MAX_BP / 2);    max = find_indel(&m, 0, ORC_MAX_BP / 2); }
    else                   { max = find_inv_tdup(&m, 1, ORC_MAX_BP / 2); max = find_inv_tdup(&m, 0, ORC_MAX_BP / 2); }  /* :111-112 */
    out->score = max;
    out->n_bp = m.n_bp;
    memcpy(out->bp, m.bp, sizeof(m.bp));
    /* K10 findAlignment(0) (:670-750): blocks only; the strings are host formatting */
    out->n_left  = left_blocks(&m, m.bp[0][0], m.bp[0][1], &out->left);
    out->n_right = right_blocks(&m, m.bp[0][2], m.bp[0][3], &out->right);
    /* printAlignment's alternative: highest-index bpoint whose alignment has any fragment (:287-306) */
    for (int i = m.n_bp - 1; i >= 1; i--) {
        orc_block *l = NULL, *r = NULL;
        int nl = left_blocks(&m, m.bp[i][0], m.bp[i][1], &l);
        int nr = right_blocks(&m, m.bp[i][2], m.bp[i][3], &r);
        if (nl == 0 && nr == 0) { free(l); free(r); continue; }
        out->alt_index = i;
        out->n_alt_left = nl; out->alt_left = l;
        out->n_alt_right = nr; out->alt_right = r;
        break;
    }
    free(rc);
    free_mat(&m);
    return 0;
}

void orc_free_result(orc_result *r)
{
    free(r->left); free(r->right); free(r->alt_left); free(r->alt_right);
    r->left = r->right = r->alt_left = r->alt_right = NULL;
}

int orc_score(const char *seq1, int len1, const char *seq2, int len2,
              int flag, int match, int mismatch, int go, int ge)
{
    if (len1 <= 0 || len2 <= 0 || !check_flag(flag)) return -1;
    mat_t m;
    if (alloc_mat(&m, len1, len2, flag)) return -2;
    char *rc = (char *)malloc((size_t)len1 + 1);
    orc_revcom(seq1, len1, rc); rc[len1] = 0;
    fill_scores(&m, seq1, rc, seq2, (short)match, (short)mismatch, (short)go, (short)ge);
    fill_maxima(&m);
    int max = 0;
    if (flag == ORC_INDEL) {
        for (int r = 0; r <= len2; r++)
            for (int c = 0; c <= len1; c++) {
                uint16_t val = (uint16_t)(AT(&m, F, r, c) + AT(&m, R, r + 1, c + 1));
                if (val > max) max = val;
            }
    } else {
        for (int r = 0; r <= len2; r++) {
            uint16_t val = (uint16_t)(AT(&m, F, r, len1) + AT(&m, R, r + 1, 1));
            if (val > max) max = val;
        }
    }
    free(rc);
    free_mat(&m);
    return max;
}
