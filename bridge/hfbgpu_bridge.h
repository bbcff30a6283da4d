/* hfbgpu_bridge.h -- HERest-side glue between HTK's in-memory structures and libhfbgpu.
 *
 * NEW source (not a modified HTK file): it only #includes HTK headers at build time.
 * The three calls map onto the seam of HTKTools/HERest.c:
 *
 *   HFBGPU_Init    after Initialise()/InitUttInfo()      (HERest.c:495-496; replaces what
 *                  InitialiseForBack hands to HFB, HFB.c:245-293)
 *   HFBGPU_FastLoad in front of LoadData() in DoForwardBackward (HERest.c:769): once the first file has shown that the
 *                  files need none of HParm's conversions, the payload of later files is read by a pool of reader
 *                  threads straight into the pinned batch buffer; FALSE = run LoadData as before
 *   HFBGPU_Queue   instead of FBFile() in DoForwardBackward (HERest.c:777): buffers the loaded
 *                  utterance; a full batch is sent to the GPU
 *   HFBGPU_Finish  after the file loop (HERest.c:534): flushes, then scatters the FP64
 *                  accumulators into TrAcc/WtAcc/MuAcc/VaAcc, hmm->hook, totalT/totalPr so that
 *                  DumpAccs (-p N) and UpdateModels (M-step) run unchanged.
 */
#ifndef HFBGPU_BRIDGE_H_
#define HFBGPU_BRIDGE_H_

#ifdef __cplusplus
extern "C" {
#endif

void HFBGPU_Init(HMMSet *hset, FBInfo *fbInfo, LogDouble pruneInit, LogDouble pruneInc,
                 LogDouble pruneLim, float minFrwdP, UPDSet uFlags, int herestTrace);
Boolean HFBGPU_FastLoad(UttInfo *utt, char *datafn, char *datafn2);
Boolean HFBGPU_Queue(FBInfo *fbInfo, UttInfo *utt, char *datafn);
void HFBGPU_Finish(int *totalT, LogDouble *totalPr);

#ifdef __cplusplus
}
#endif
#endif
