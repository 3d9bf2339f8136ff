// dev probe: statistics of the reference parse on config-2/3/4 blocks (probes per sequence, offsets, lengths)
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline uint32_t ld32(const uint8_t*p){uint32_t v;memcpy(&v,p,4);return v;}
static inline uint64_t ld64(const uint8_t*p){uint64_t v;memcpy(&v,p,8);return v;}
static inline uint32_t slot5(const uint8_t*p){return (uint32_t)(((ld64(p)<<24)*889523592379ull)>>52);}
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb");fseek(f,0,SEEK_END);size_t fl=ftell(f);fseek(f,0,SEEK_SET);
  uint8_t*file=malloc(fl);fread(file,1,fl,f);fclose(f);
  size_t bs=atoi(argv[2]); int nb=atoi(argv[3]);
  uint8_t*in=malloc(bs+8);
  uint64_t nseq=0,nprobe=0,probe_hist[40]={0},off_hist[20]={0},ml_hist[40]={0},lit_hist[40]={0},coll=0,steps4=0,steps2=0,steps8=0;
  uint64_t back=0, far_sector=0;
  for(int b=0;b<nb;b++){
    for(size_t i=0;i<bs;i++) in[i]=file[((size_t)b*bs+i)%fl];
    uint32_t tab[4096];memset(tab,0,sizeof tab);
    size_t n=bs,anchor=0,cur=0,last_probe=n-12;
    tab[slot5(in)]=0;cur=1;
    for(;;){
      size_t misses=32,next=cur,cand;int np=0;
      for(;;){size_t step=misses>>5;misses++;cur=next;next+=step;
        if(cur>last_probe) goto done;
        uint32_t s=slot5(in+cur);cand=tab[s];tab[s]=cur;np++;
        if(cur-cand>65535)continue;
        if(ld32(in+cand)==ld32(in+cur))break;}
      nprobe+=np;probe_hist[np<39?np:39]++;
      steps4+=(np+3)/4;steps2+=(np+1)/2;steps8+=(np+7)/8;
      size_t c0=cur;
      while(cand>0&&cur>anchor&&in[cur-1]==in[cand-1]){cur--;cand--;}
      back+=c0-cur;
      size_t lit=cur-anchor;uint32_t dist=cur-cand;
      cur+=4;cand+=4;size_t e=0;while(cur<n-6&&in[cur]==in[cand]){cur++;cand++;e++;}
      tab[slot5(in+cur-2)]=cur-2;
      nseq++;
      int lg=0;while((1u<<lg)<dist)lg++;off_hist[lg]++;
      ml_hist[(e+4)/8<39?(e+4)/8:39]++;lit_hist[lit<39?lit:39]++;
      anchor=cur;
    }
    done:;
  }
  printf("blocks %d seq/block %.1f probes/seq %.3f steps(4-wide)/seq %.3f steps2 %.3f steps8 %.3f backtrack/seq %.3f\n",nb,(double)nseq/nb,(double)nprobe/nseq,(double)steps4/nseq,(double)steps2/nseq,(double)steps8/nseq,(double)back/nseq);
  printf("probes hist:");for(int i=1;i<40;i++)printf(" %d:%.3f",i,(double)probe_hist[i]/nseq);printf("\n");
  printf("offset log2 hist:");for(int i=0;i<=16;i++)printf(" %d:%.3f",i,(double)off_hist[i]/nseq);printf("\n");
  printf("matchlen/8 hist:");for(int i=0;i<40;i++)printf(" %d:%.3f",i*8,(double)ml_hist[i]/nseq);printf("\n");
  printf("lit hist:");for(int i=0;i<40;i++)printf(" %d:%.3f",i,(double)lit_hist[i]/nseq);printf("\n");
  return 0;}
