# Times the REAL genomicMateSelectR::predCrossVars (all host cores) on the synthetic cassava-scale inputs of
# SURVEY.md section 8d, for anyone who has R. Not executable in this repository's image (no R there): bench.py's
# --impl reference / cpu_baseline legs time a C+OpenMP restatement of the same algorithm instead.
#   Rscript ref_predCrossVars.R [n_crosses = 8] [n_snps_per_chr = 1833] [n_chr = 18]
suppressMessages({ library(genomicMateSelectR); library(tibble) })
args <- as.numeric(commandArgs(trailingOnly = TRUE))
ncross <- if (length(args) >= 1) args[1] else 8
mc <- if (length(args) >= 2) args[2] else 1833
nchr <- if (length(args) >= 3) args[3] else 18
P <- 1000; ntrait <- 5
set.seed(20261019)
m <- mc * nchr
snps <- paste0(rep(seq_len(nchr), each = mc), "_", seq_len(m))
genmap <- unlist(lapply(seq_len(nchr), function(c) sort(runif(mc, 0, 130)))); names(genmap) <- snps
recombFreqMat <- 1 - (2 * genmap2recombfreq(genmap, nChr = nchr))
p <- runif(m, 0.05, 0.95)
gids <- as.character(seq_len(P))
haploMat <- matrix(as.integer(runif(2 * P * m) < rep(p, each = 2 * P)), nrow = 2 * P,
                   dimnames = list(as.vector(rbind(paste0(gids, "_HapA"), paste0(gids, "_HapB"))), snps))
eff <- function(sd) setNames(lapply(seq_len(ntrait), function(t) matrix(rnorm(m, 0, sd), nrow = 1, dimnames = list(NULL, snps))),
                             paste0("Trait", seq_len(ntrait)))
AddEffectList <- eff(1); DomEffectList <- eff(0.5)
crosses <- tibble(sireID = sample(gids, ncross, TRUE), damID = sample(gids, ncross, TRUE))
ncores <- parallel::detectCores()
t0 <- proc.time()[3]
res <- predCrossVars(CrossesToPredict = crosses, modelType = "AD", AddEffectList = AddEffectList,
                     DomEffectList = DomEffectList, haploMat = haploMat, recombFreqMat = recombFreqMat,
                     ncores = min(ncores, ncross), nBLASthreads = 1)
dt <- proc.time()[3] - t0
cat(sprintf('{"impl": "reference-R", "crosses": %d, "seconds": %.2f, "crosses_per_s": %.4f, "cores": %d}\n',
            ncross, dt, ncross / dt, ncores))
