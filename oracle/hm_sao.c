/* TEST INFRASTRUCTURE - CPU oracle of HM 16.14's SAO statistics (second stage of the per-picture
 * tail): TEncSampleAdaptiveOffset::getStatistics / getBlkStats
 * (/root/reference/source/Lib/TLibEncoder/TEncSampleAdaptiveOffset.cpp:287-332, :911-1290) restated loop for loop - the
 * same sign line buffers, the same first-line special cases - for the encoder's default set-up (statistics from the
 * deblocked picture only: isCalculatePreDeblockSamples = false, skip lines 5 / 4 luma and 3 / 2 chroma, :124-131).  The
 * checker of fhc_deblock_pictures_stats, whose kernel computes the same numbers sample by sample from a closed-form
 * membership test; nothing in the product may include or link it.  PINNED against the reference's own m_statData
 * (oracle/_ref/TAppEncoderDump with FHC_SAO_DUMP, tests/test_deblock.py). */
#include <stdint.h>
#include <string.h>

typedef struct { int64_t diff[32]; int64_t count[32]; } hm_sao_stat;

static int sgn(int v) { return (v > 0) - (v < 0); }

static void blk_stats(int comp, int bd, hm_sao_stat* st, const int16_t* srcBlk, const int16_t* orgBlk, int srcStride, int orgStride, int width, int height,
                      int isLeftAvail, int isRightAvail, int isAboveAvail, int isBelowAvail, int isAboveLeftAvail, int isAboveRightAvail)
{
  int8_t buf1[66], buf2[66];
  const int skipR = comp ? 3 : 5, skipB = comp ? 2 : 4;
  int typeIdx;
  for (typeIdx = 0; typeIdx < 5; typeIdx++)
  {
    hm_sao_stat* s = &st[typeIdx];
    const int16_t* srcLine = srcBlk;
    const int16_t* orgLine = orgBlk;
    int64_t* diff = s->diff;
    int64_t* count = s->count;
    int x, y, startX, startY, endX, endY, edgeType, firstLineStartX, firstLineEndX;
    int8_t signLeft, signRight, signDown;
    memset(s, 0, sizeof(*s));
    switch (typeIdx)
    {
    case 0:                                                                           /* SAO_TYPE_EO_0 :950 */
      diff += 2; count += 2;
      endY = isBelowAvail ? (height - skipB) : height;
      startX = isLeftAvail ? 0 : 1;
      endX = isRightAvail ? (width - skipR) : (width - 1);
      for (y = 0; y < endY; y++)
      {
        signLeft = (int8_t)sgn(srcLine[startX] - srcLine[startX - 1]);
        for (x = startX; x < endX; x++)
        {
          signRight = (int8_t)sgn(srcLine[x] - srcLine[x + 1]);
          edgeType = signRight + signLeft;
          signLeft = (int8_t)-signRight;
          diff[edgeType] += (orgLine[x] - srcLine[x]);
          count[edgeType]++;
        }
        srcLine += srcStride; orgLine += orgStride;
      }
      break;
    case 1:                                                                           /* SAO_TYPE_EO_90 :1008 */
    {
      int8_t* signUpLine = buf1;
      const int16_t* srcLineAbove;
      diff += 2; count += 2;
      startX = 0;
      startY = isAboveAvail ? 0 : 1;
      endX = isRightAvail ? (width - skipR) : width;
      endY = isBelowAvail ? (height - skipB) : (height - 1);
      if (!isAboveAvail) { srcLine += srcStride; orgLine += orgStride; }
      srcLineAbove = srcLine - srcStride;
      for (x = startX; x < endX; x++) signUpLine[x] = (int8_t)sgn(srcLine[x] - srcLineAbove[x]);
      for (y = startY; y < endY; y++)
      {
        const int16_t* srcLineBelow = srcLine + srcStride;
        for (x = startX; x < endX; x++)
        {
          signDown = (int8_t)sgn(srcLine[x] - srcLineBelow[x]);
          edgeType = signDown + signUpLine[x];
          signUpLine[x] = (int8_t)-signDown;
          diff[edgeType] += (orgLine[x] - srcLine[x]);
          count[edgeType]++;
        }
        srcLine += srcStride; orgLine += orgStride;
      }
      break;
    }
    case 2:                                                                           /* SAO_TYPE_EO_135 :1076 */
    {
      int8_t *signUpLine = buf1, *signDownLine = buf2, *signTmpLine;
      const int16_t *srcLineBelow, *srcLineAbove;
      diff += 2; count += 2;
      startX = isLeftAvail ? 0 : 1;
      endX = isRightAvail ? (width - skipR) : (width - 1);
      endY = isBelowAvail ? (height - skipB) : (height - 1);
      srcLineBelow = srcLine + srcStride;
      for (x = startX; x < endX + 1; x++) signUpLine[x] = (int8_t)sgn(srcLineBelow[x] - srcLine[x - 1]);
      srcLineAbove = srcLine - srcStride;
      firstLineStartX = isAboveLeftAvail ? 0 : 1;
      firstLineEndX = isAboveAvail ? endX : 1;
      for (x = firstLineStartX; x < firstLineEndX; x++)
      {
        edgeType = sgn(srcLine[x] - srcLineAbove[x - 1]) - signUpLine[x + 1];
        diff[edgeType] += (orgLine[x] - srcLine[x]);
        count[edgeType]++;
      }
      srcLine += srcStride; orgLine += orgStride;
      for (y = 1; y < endY; y++)
      {
        srcLineBelow = srcLine + srcStride;
        for (x = startX; x < endX; x++)
        {
          signDown = (int8_t)sgn(srcLine[x] - srcLineBelow[x + 1]);
          edgeType = signDown + signUpLine[x];
          diff[edgeType] += (orgLine[x] - srcLine[x]);
          count[edgeType]++;
          signDownLine[x + 1] = (int8_t)-signDown;
        }
        signDownLine[startX] = (int8_t)sgn(srcLineBelow[startX] - srcLine[startX - 1]);
        signTmpLine = signUpLine; signUpLine = signDownLine; signDownLine = signTmpLine;
        srcLine += srcStride; orgLine += orgStride;
      }
      break;
    }
    case 3:                                                                           /* SAO_TYPE_EO_45 :1160 */
    {
      int8_t* signUpLine = buf1 + 1;
      const int16_t *srcLineBelow, *srcLineAbove;
      diff += 2; count += 2;
      startX = isLeftAvail ? 0 : 1;
      endX = isRightAvail ? (width - skipR) : (width - 1);
      endY = isBelowAvail ? (height - skipB) : (height - 1);
      srcLineBelow = srcLine + srcStride;
      for (x = startX - 1; x < endX; x++) signUpLine[x] = (int8_t)sgn(srcLineBelow[x] - srcLine[x + 1]);
      srcLineAbove = srcLine - srcStride;
      firstLineStartX = isAboveAvail ? startX : endX;
      firstLineEndX = (!isRightAvail && isAboveRightAvail) ? width : endX;
      for (x = firstLineStartX; x < firstLineEndX; x++)
      {
        edgeType = sgn(srcLine[x] - srcLineAbove[x + 1]) - signUpLine[x - 1];
        diff[edgeType] += (orgLine[x] - srcLine[x]);
        count[edgeType]++;
      }
      srcLine += srcStride; orgLine += orgStride;
      for (y = 1; y < endY; y++)
      {
        srcLineBelow = srcLine + srcStride;
        for (x = startX; x < endX; x++)
        {
          signDown = (int8_t)sgn(srcLine[x] - srcLineBelow[x - 1]);
          edgeType = signDown + signUpLine[x];
          diff[edgeType] += (orgLine[x] - srcLine[x]);
          count[edgeType]++;
          signUpLine[x - 1] = (int8_t)-signDown;
        }
        signUpLine[endX - 1] = (int8_t)sgn(srcLineBelow[endX - 1] - srcLine[endX]);
        srcLine += srcStride; orgLine += orgStride;
      }
      break;
    }
    default:                                                                          /* SAO_TYPE_BO :1246 */
    {
      const int shiftBits = bd - 5;
      startX = 0;
      endX = isRightAvail ? (width - skipR) : width;
      endY = isBelowAvail ? (height - skipB) : height;
      for (y = 0; y < endY; y++)
      {
        for (x = startX; x < endX; x++)
        {
          const int bandIdx = srcLine[x] >> shiftBits;
          diff[bandIdx] += (orgLine[x] - srcLine[x]);
          count[bandIdx]++;
        }
        srcLine += srcStride; orgLine += orgStride;
      }
    }
    }
  }
}

/* getStatistics :287.  src: deblocked planes (int16, Y [H][W], Cb, Cr [H/2][W/2]); org: original planes, same layout;
 * stats: [numCtus][3][5].  Availability = picture boundaries (one slice, one tile: deriveLoopFilterBoundaryAvailibility
 * reduces to them; the right / below flags are overridden to picture boundaries by getStatistics itself, :304-308). */
int hm_sao_statistics(int width, int height, int bitDepth, const int16_t* const src[3], const int16_t* const org[3], hm_sao_stat* stats)
{
  const int wc = (width + 63) / 64, hc = (height + 63) / 64;
  int addr, comp;
  for (addr = 0; addr < wc * hc; addr++)
  {
    const int yPos = (addr / wc) * 64, xPos = (addr % wc) * 64;
    const int h = yPos + 64 > height ? height - yPos : 64, w = xPos + 64 > width ? width - xPos : 64;
    const int left = xPos > 0, above = yPos > 0, right = xPos + 64 < width, below = yPos + 64 < height;
    for (comp = 0; comp < 3; comp++)
    {
      const int sh = comp ? 1 : 0, stride = width >> sh;
      blk_stats(comp, bitDepth, &stats[(addr * 3 + comp) * 5], src[comp] + (yPos >> sh) * stride + (xPos >> sh), org[comp] + (yPos >> sh) * stride + (xPos >> sh),
                stride, stride, w >> sh, h >> sh, left, right, above, below, left && above, above && right);
    }
  }
  return 0;
}
