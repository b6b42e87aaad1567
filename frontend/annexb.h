// annexb.h — byte-stream helpers shared by the harnesses (hmdec_cli, hmdec_mt)
#ifndef HM_ANNEXB_H
#define HM_ANNEXB_H
#include <cstdio>
#include <cstdint>
#include <vector>
#include <utility>

static inline bool readFile(const char* path, std::vector<uint8_t>& out)
{
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  out.resize(n);
  bool ok = fread(out.data(), 1, n, f) == (size_t)n;
  fclose(f);
  return ok;
}

// Annex B (B.2): NAL units are delimited by 00 00 01; trailing zero bytes belong to the delimiter.
static inline void splitAnnexB(const std::vector<uint8_t>& s, std::vector<std::pair<size_t, size_t> >& nals)
{
  size_t n = s.size(), i = 0, start = (size_t)-1;
  while (i + 2 < n)
  {
    if (s[i] == 0 && s[i + 1] == 0 && s[i + 2] == 1)
    {
      if (start != (size_t)-1)
      {
        size_t end = i;
        while (end > start && s[end - 1] == 0) end--;
        nals.push_back(std::make_pair(start, end - start));
      }
      start = i + 3;
      i += 3;
    }
    else i++;
  }
  if (start != (size_t)-1 && start < n)
  {
    size_t end = n;
    while (end > start && s[end - 1] == 0) end--;
    nals.push_back(std::make_pair(start, end - start));
  }
}


#endif
