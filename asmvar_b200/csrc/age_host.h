// age_host.h -- internal declarations shared by the host-side translation units.
#ifndef AGE_HOST_H
#define AGE_HOST_H
#include "age_b200.h"
#include <cstddef>

namespace age {

char complement(char c);                 // Sequence::complement  (Sequence.hh:138-150)
bool same_nuc(char a, char b);           // Sequence::sameNuc     (Sequence.hh:182-188)

size_t format_alignment(const age_seqview *s1, const age_seqview *s2, const age_job *job,
                        const age_result *res, char *buf, size_t cap);

} // namespace age
#endif
