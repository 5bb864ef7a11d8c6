/* fastga_b200_hook.c -- the reference-side binding of libfastga_b200.so.
 *
 * #included into the reference's FastGA.c (just above main) by integration/FastGA_b200.patch when it is
 * compiled with -DFASTGA_B200; the patch also skips the two system("GIXmake ...") calls and calls
 * b200_main() right after the sources have been resolved to GDBs.  From there the B200 library
 * does what FastGA.c:4867-5205 does on the CPU (GIX build of both genomes, adaptamer merge, seed
 * sort, chain search, Local_Alignment, redundancy filter, la_sort order); this file only moves
 * data between the reference's own structures and the C-ABI:
 *   in : the GDBs as Read_GDB leaves them (GDB.h:28-72) and the .bps images,
 *   out: Overlap records through the reference's own .1aln writer (alncode.c:272-305), then the
 *        reference's own ALNtoPAF / ALNtoPSL when -paf / -psl was asked for (FastGA.c:5205-5250),
 *        and the -v lines scripts parse (FastGA.c:2485, :4358-4389).
 * Everything static in FastGA.c (options, paths, Prog_Name ...) is visible here.
 */
#include "fastga_b200.h"

static uint8 *b200_bps_image(GDB *gdb, int64 *nbytes)
{ int64  n, got;
  int    i;
  uint8 *img;
  FILE  *f;

  n = 0;
  for (i = 0; i < gdb->ncontig; i++)
    if (gdb->contigs[i].boff >= 0)
      { int64 e = gdb->contigs[i].boff + COMPRESSED_LEN(gdb->contigs[i].clen);
        if (e > n)
          n = e;
      }
  img = Malloc(n+16,"Allocating .bps image");
  if (img == NULL)
    Clean_Exit(1);
  if (gdb->seqstate == EXTERNAL)
    { f = (FILE *) gdb->seqs;
      rewind(f);
      got = fread(img,1,n,f);
      if (got != n)
        { fprintf(stderr,"%s: Cannot read %lld bytes of %s\n",Prog_Name,n,gdb->seqpath);
          Clean_Exit(1);
        }
    }
  else if (gdb->seqstate == COMPRESSED)
    memcpy(img,gdb->seqs,n);
  else
    { fprintf(stderr,"%s: GDB sequence state %d not supported by the B200 path\n",Prog_Name,gdb->seqstate);
      Clean_Exit(1);
    }
  *nbytes = n;
  return (img);
}

static void b200_contig_table(GDB *gdb, int64 **clen, int64 **boff)
{ int i;

  *clen = Malloc(sizeof(int64)*gdb->ncontig,"Allocating contig table");
  *boff = Malloc(sizeof(int64)*gdb->ncontig,"Allocating contig table");
  if (*clen == NULL || *boff == NULL)
    Clean_Exit(1);
  for (i = 0; i < gdb->ncontig; i++)
    { (*clen)[i] = gdb->contigs[i].clen;
      (*boff)[i] = gdb->contigs[i].boff;
    }
}

static char *b200_gdb_extn(char *path, char *root)       /* .gdb if present else .1gdb (FastGA.c:4918-4926) */
{ FILE *f = fopen(Catenate(path,"/",root,".gdb"),"r");
  if (f == NULL)
    return (".1gdb");
  fclose(f);
  return (".gdb");
}

static void b200_main(GDB *gdb1, GDB *gdb2)
{ fgb_alns     *alns;
  fgb_run_stats st;
  int64         n1, n2, nrec, *clen1, *boff1, *clen2, *boff2;
  uint8        *bps1, *bps2;
  int           rc;

  if (VERBOSE || LOG_FILE)
    StartTime();

  GEXTN1 = b200_gdb_extn(PATH1,ROOT1);
  if (Read_GDB(gdb1,Catenate(PATH1,"/",ROOT1,GEXTN1)) < 0)
    Clean_Exit(1);
  if (SELF)
    gdb2 = gdb1;
  else
    { GEXTN2 = b200_gdb_extn(PATH2,ROOT2);
      if (Read_GDB(gdb2,Catenate(PATH2,"/",ROOT2,GEXTN2)) < 0)
        Clean_Exit(1);
    }

  if (OUT_TYPE != 2)                                       /* temporary .1aln for the converter (FastGA.c:4946-4951) */
    { ONE_ROOT = Strdup(Numbered_Suffix("_oaln.",getpid(),""),"Allocating temp name");
      ONE_PATH = SORT_PATH;
    }

  bps1 = b200_bps_image(gdb1,&n1);
  b200_contig_table(gdb1,&clen1,&boff1);
  if (VERBOSE)
    { fprintf(stderr,"\n  B200: GIX build, adaptive seed merge, seed sort and alignment search on the device\n");
      fflush(stderr);
    }
  if (SELF)
    rc = fgb_fastga_self(bps1,n1,gdb1->ncontig,clen1,boff1,gdb1->freq,
                         FREQ,CHAIN_BREAK,CHAIN_MIN,ALIGN_MIN,ALIGN_RATE,&alns,&st,NULL);
  else
    { bps2 = b200_bps_image(gdb2,&n2);
      b200_contig_table(gdb2,&clen2,&boff2);
      rc = fgb_fastga(bps1,n1,gdb1->ncontig,clen1,boff1,gdb1->freq,
                      bps2,n2,gdb2->ncontig,clen2,boff2,
                      FREQ,CHAIN_BREAK,CHAIN_MIN,ALIGN_MIN,ALIGN_RATE,&alns,&st,NULL);
      free(bps2); free(clen2); free(boff2);
    }
  free(bps1); free(clen1); free(boff1);
  if (rc != 0)
    { fprintf(stderr,"%s: fastga_b200 failed with code %d\n",Prog_Name,rc);
      Clean_Exit(1);
    }

  nrec = fgb_alns_count(alns);
  { int   *fld  = Malloc(sizeof(int)*9*(nrec+1),"Allocating records");
    int64 *toff = Malloc(sizeof(int64)*(nrec+1),"Allocating records");
    uint8 *pool = Malloc(fgb_alns_pool_bytes(alns)+16,"Allocating records");
    int64 *trace64, k, tmax, ncov;
    char  *db1_name, *db2_name, *cpath;
    OneFile *of;

    if (fld == NULL || toff == NULL || pool == NULL)
      Clean_Exit(1);
    fgb_alns_get(alns,fld,toff,pool);

    tmax = 2;
    ncov = 0;
    for (k = 0; k < nrec; k++)
      { if (fld[9*k+8] > tmax)
          tmax = fld[9*k+8];
        ncov += fld[9*k+5] - fld[9*k+3];
      }
    trace64 = Malloc(sizeof(int64)*tmax,"Allocating int64 trace vector");

    if (VERBOSE)
      { if (st.nseeds > 0)
          fprintf(stderr,"\n  Total seeds = %lld, ave. len = %.1f, seeds per genome position = %.1f\n",
                         st.nseeds,(1.*st.sumlen)/st.nseeds,(1.*st.nseeds)/gdb1->seqtot);
        if (nrec == 0)
          fprintf(stderr,"\n  Total hits over %dbp = %lld, %lld aln's, 0 %s\n",
                         CHAIN_MIN/2,st.nhits,fgb_alns_raw_count(alns),"non-redundant aln's of ave len 0");
        else
          fprintf(stderr,"\n  Total hits over %dbp = %lld, %lld aln's, %lld %s %lld\n",
                         CHAIN_MIN/2,st.nhits,fgb_alns_raw_count(alns),nrec,"non-redundant aln's of ave len",ncov/nrec);
        fflush(stderr);
      }

    /* the header la_merge writes (FastGA.c:4049-4073) */
    if (TYPE1 < IS_GDB && !KEEP)
      db1_name = Strdup(SPATH1,"db1_name");
    else
      db1_name = Strdup(Catenate(PATH1,"/",ROOT1,GEXTN1),"db1_name");
    if (SELF)
      db2_name = NULL;
    else if (TYPE2 < IS_GDB && !KEEP)
      db2_name = Strdup(SPATH2,"db2_name");
    else
      db2_name = Strdup(Catenate(PATH2,"/",ROOT2,GEXTN2),"db2_name");
    cpath = getcwd(NULL,0);
    of = open_Aln_Write(Catenate(ONE_PATH,"/",ONE_ROOT,".1aln"),1,
                        Prog_Name,VERSION,Command_Line,TSPACE,db1_name,db2_name,cpath);
    Write_Skeleton(of,gdb1);
    if (!SELF)
      Write_Skeleton(of,gdb2);
    free(cpath);
    free(db2_name);
    free(db1_name);

    for (k = 0; k < nrec; k++)                              /* records arrive in la_sort order */
      { Overlap ov;
        int    *r = fld + 9*k;

        ov.flags = r[0] ? COMP_FLAG : 0;
        ov.aread = r[1];
        ov.bread = r[2];
        ov.path.abpos = r[3];
        ov.path.bbpos = r[4];
        ov.path.aepos = r[5];
        ov.path.bepos = r[6];
        ov.path.diffs = r[7];
        ov.path.tlen  = r[8];
        ov.path.trace = pool + toff[k];
        Write_Aln_Overlap(of,&ov);
        Write_Aln_Trace(of,pool + toff[k],r[8],trace64,0);
      }
    oneFileClose(of);
    free(trace64); free(pool); free(toff); free(fld);
  }
  fgb_alns_free(alns);

  if (VERBOSE)
    TimeTo(stderr,0,LOG_FILE==NULL);

  if (OUT_TYPE != 2)                                       /* PAF / PSL through the reference's own converters */
    { char *command = Malloc(strlen(ONE_ROOT)+strlen(ONE_PATH)+100,"Allocating command buffer");
      if (command == NULL)
        Clean_Exit(1);
      if (OUT_TYPE == 0)
        sprintf(command,"ALNtoPAF %s %s -T%d %s/%s",
                        (OUT_OPT&PAFM)? "-m" : ((OUT_OPT&PAFX)? "-x" : ""),
                        (OUT_OPT&PAFS)? "-s" : ((OUT_OPT&PAFL)? "-S" : ""),
                        NTHREADS,ONE_PATH,ONE_ROOT);
      else
        sprintf(command,"ALNtoPSL -T%d %s/%s",NTHREADS,ONE_PATH,ONE_ROOT);
      rc = system(command);
      unlink(Catenate(ONE_PATH,"/",ONE_ROOT,".1aln"));
      if (rc != 0)
        { fprintf(stderr,"\n%s: Call to %s failed\n",Prog_Name,OUT_TYPE == 0 ? "ALNtoPAF" : "ALNtoPSL");
          Clean_Exit(1);
        }
      free(command);
    }

  if (VERBOSE)
    TimeTo(stderr,1,0);
  if (!SELF)
    Close_GDB(gdb2);
  Close_GDB(gdb1);
  exit (0);
}
