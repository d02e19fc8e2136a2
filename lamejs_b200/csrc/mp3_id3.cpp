/* mp3_id3.cpp -- ID3 version 1 / 1.1 and version 2.3 tags, host side.
 *
 * lamejs itself carries only a stub (`function ID3Tag()` in src/js/index.js:56-64 and the empty src/js/ID3TagSpec.js) and
 * switches the automatic tags off (index.js:109); the writer lives in the Java original, src/main/java/mp3/ID3Tag.java, and
 * that is what is restated here: id3tag_init :209-215, the id3tag_set_* family :560-740, lame_get_id3v2_tag :961-1102,
 * lame_get_id3v1_tag :1141-1189, the frame writers set_frame_custom2 :872-892 and set_frame_comment :834-870.
 * No engine in the build image runs Java, so this row is checked by an independent reader of the two published formats
 * (tests/test_id3.py), not against the reference's own output.
 * Not restated: UCS-2 text (id3v2_add_ucs2), album art (APIC), free-form `--tv` frames, and the "sloppy" genre-name matcher
 * (ID3Tag.java:657-698) -- a genre is a number 0..147 or one of the 148 names spelled as in the table, anything else is
 * "Other" in the version 1 tag and the caller's text in the version 2 tag, as the Java does when its matcher finds nothing.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/mp3b200.h"

namespace {

enum { CHANGED_FLAG = 1, ADD_V2_FLAG = 2, V1_ONLY_FLAG = 4, V2_ONLY_FLAG = 8, SPACE_V1_FLAG = 16, PAD_V2_FLAG = 32 };

const char* const kGenres[148] = {
    "Blues", "Classic Rock", "Country", "Dance", "Disco", "Funk", "Grunge", "Hip-Hop", "Jazz", "Metal", "New Age", "Oldies", "Other", "Pop",
    "R&B", "Rap", "Reggae", "Rock", "Techno", "Industrial", "Alternative", "Ska", "Death Metal", "Pranks", "Soundtrack", "Euro-Techno",
    "Ambient", "Trip-Hop", "Vocal", "Jazz+Funk", "Fusion", "Trance", "Classical", "Instrumental", "Acid", "House", "Game", "Sound Clip",
    "Gospel", "Noise", "Alternative Rock", "Bass", "Soul", "Punk", "Space", "Meditative", "Instrumental Pop", "Instrumental Rock", "Ethnic",
    "Gothic", "Darkwave", "Techno-Industrial", "Electronic", "Pop-Folk", "Eurodance", "Dream", "Southern Rock", "Comedy", "Cult", "Gangsta",
    "Top 40", "Christian Rap", "Pop/Funk", "Jungle", "Native US", "Cabaret", "New Wave", "Psychedelic", "Rave", "Showtunes", "Trailer", "Lo-Fi",
    "Tribal", "Acid Punk", "Acid Jazz", "Polka", "Retro", "Musical", "Rock & Roll", "Hard Rock", "Folk", "Folk-Rock", "National Folk", "Swing",
    "Fast Fusion", "Bebob", "Latin", "Revival", "Celtic", "Bluegrass", "Avantgarde", "Gothic Rock", "Progressive Rock", "Psychedelic Rock",
    "Symphonic Rock", "Slow Rock", "Big Band", "Chorus", "Easy Listening", "Acoustic", "Humour", "Speech", "Chanson", "Opera", "Chamber Music",
    "Sonata", "Symphony", "Booty Bass", "Primus", "Porn Groove", "Satire", "Slow Jam", "Club", "Tango", "Samba", "Folklore", "Ballad",
    "Power Ballad", "Rhythmic Soul", "Freestyle", "Duet", "Punk Rock", "Drum Solo", "A Cappella", "Euro-House", "Dance Hall", "Goa",
    "Drum & Bass", "Club-House", "Hardcore", "Terror", "Indie", "BritPop", "Negerpunk", "Polsk Punk", "Beat", "Christian Gangsta", "Heavy Metal",
    "Black Metal", "Crossover", "Contemporary Christian", "Christian Rock", "Merengue", "Salsa", "Thrash Metal", "Anime", "JPop", "SynthPop"};

struct Frame { char id[5]; bool comment; std::string text; };

/* the state id3tag_init + the id3tag_set_* calls leave in gfc.tag_spec */
struct Spec {
  int flags = 0, year = 0, track = 0, genre = 255, padding = 128;
  std::string title, artist, album, comment;
  std::vector<Frame> frames;                     /* the version 2 frame list, in the order the frames were first set */

  void set_frame(const char* id, const std::string& text, bool comment = false) {   /* id3v2_add_latin1 (single-instance frames) */
    for (auto& f : frames)
      if (!strcmp(f.id, id)) { f.text = text; return; }
    Frame f;
    memcpy(f.id, id, 5); f.comment = comment; f.text = text;
    frames.push_back(f);
  }
  /* copyV1ToV2: the frame is recorded, the flags stay as they were */
  void mirror(const char* id, const std::string& text) { set_frame(id, text); }
};

bool present(const char* s) { return s && s[0]; }

/* Integer.parseInt / Integer.valueOf on a decimal string; false = NumberFormatException */
bool parse_int(const std::string& s, long* out) {
  if (s.empty()) return false;
  char* end = nullptr;
  const long v = strtol(s.c_str(), &end, 10);
  if (*end != 0 || s[0] == ' ') return false;
  *out = v;
  return true;
}

int build_spec(const mp3b200_id3tag* t, Spec* sp) {
  /* id3tag_init: the encoder frame comes first */
  sp->mirror("TSSE", "LAME 32bits version 3.98.4 (http://www.mp3dev.org/)");
  if (!t) return 0;
  if (present(t->title)) { sp->title = t->title; sp->flags |= CHANGED_FLAG; sp->mirror("TIT2", t->title); }
  if (present(t->artist)) { sp->artist = t->artist; sp->flags |= CHANGED_FLAG; sp->mirror("TPE1", t->artist); }
  if (present(t->album)) { sp->album = t->album; sp->flags |= CHANGED_FLAG; sp->mirror("TALB", t->album); }
  if (present(t->year)) {
    long num = 0;
    if (!parse_int(t->year, &num)) return MP3B200_ERR_CONFIG;
    if (num < 0) num = 0;
    if (num > 9999) num = 9999;
    if (num != 0) { sp->year = (int)num; sp->flags |= CHANGED_FLAG; }
    sp->mirror("TYER", t->year);
  }
  if (present(t->comment)) { sp->comment = t->comment; sp->flags |= CHANGED_FLAG; sp->set_frame("COMM", t->comment, true); }
  if (present(t->track)) {
    const std::string tr = t->track;
    const size_t slash = tr.find('/');
    long num = 0;
    if (!parse_int(slash == std::string::npos ? tr : tr.substr(0, slash), &num)) return MP3B200_ERR_CONFIG;
    if (num < 1 || num > 255) { num = 0; sp->flags |= CHANGED_FLAG | ADD_V2_FLAG; }   /* out of the version 1 range */
    if (num != 0) { sp->track = (int)num; sp->flags |= CHANGED_FLAG; }
    if (slash != std::string::npos) sp->flags |= CHANGED_FLAG | ADD_V2_FLAG;
    sp->mirror("TRCK", tr);
  }
  if (present(t->genre)) {
    std::string g = t->genre;
    long num = 0;
    bool unknown = false;
    if (parse_int(g, &num)) {
      if (num < 0 || num >= 148) return MP3B200_ERR_CONFIG;
      g = kGenres[num];
    } else {
      num = 148;
      for (int i = 0; i < 148; i++) if (g == kGenres[i]) { num = i; break; }
      if (num == 148) { num = 12; unknown = true; }                                    /* GENRE_INDEX_OTHER */
    }
    sp->genre = (int)num;
    sp->flags |= CHANGED_FLAG;
    if (unknown) sp->flags |= ADD_V2_FLAG;
    sp->mirror("TCON", g);
  }
  /* the id3tag_add_v2 / v1_only / v2_only / space_v1 / set_pad switches, applied in that order */
  const int f = t->flags;
  if (f & MP3B200_ID3_ADD_V2) { sp->flags &= ~V1_ONLY_FLAG; sp->flags |= ADD_V2_FLAG; }
  if (f & MP3B200_ID3_V1_ONLY) { sp->flags &= ~(ADD_V2_FLAG | V2_ONLY_FLAG); sp->flags |= V1_ONLY_FLAG; }
  if (f & MP3B200_ID3_V2_ONLY) { sp->flags &= ~V1_ONLY_FLAG; sp->flags |= V2_ONLY_FLAG; }
  if (f & MP3B200_ID3_SPACE_V1) { sp->flags &= ~V2_ONLY_FLAG; sp->flags |= SPACE_V1_FLAG; }
  if (f & MP3B200_ID3_PAD_V2) {
    sp->flags &= ~V1_ONLY_FLAG;
    sp->flags |= PAD_V2_FLAG | ADD_V2_FLAG;
    sp->padding = t->padding > 0 ? t->padding : 128;
  }
  return 0;
}

size_t frame_size(const Frame& f) { return f.comment ? 10 + 1 + 3 + 1 + f.text.size() : 10 + 1 + f.text.size(); }

uint8_t* put32(uint8_t* p, unsigned long v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; return p + 4; }

uint8_t* text_field(uint8_t* p, const std::string* text, int size, int pad) {   /* set_text_field */
  for (int i = 0; i < size; i++) *p++ = (text && (size_t)i < text->size()) ? (uint8_t)(*text)[i] : (uint8_t)pad;
  return p;
}

}  // namespace

extern "C" {

int mp3b200_id3v2_tag(const mp3b200_id3tag* t, uint8_t* buf, int cap) {
  Spec sp;
  const int rc = build_spec(t, &sp);
  if (rc) return rc;
  if (sp.flags & V1_ONLY_FLAG) return 0;
  /* written if asked for, or if a field does not fit the version 1 tag */
  if (!((sp.flags & (ADD_V2_FLAG | V2_ONLY_FLAG)) || sp.title.size() > 30 || sp.artist.size() > 30 || sp.album.size() > 30 ||
        sp.comment.size() > 30 || (sp.track != 0 && sp.comment.size() > 28)))
    return 0;
  if (t && t->num_samples >= 0 && t->samplerate > 0) {       /* id3v2AddAudioDuration: play length in ms */
    double ms = (double)t->num_samples;
    ms *= 1000;
    ms /= t->samplerate;
    long long playlength = ms > 2147483647.0 ? 2147483647LL : ms < 0 ? 0 : (long long)ms;
    char b[32];
    snprintf(b, sizeof b, "%lld", playlength);
    sp.mirror("TLEN", b);
  }
  size_t tag_size = 10;
  for (const auto& f : sp.frames) tag_size += frame_size(f);
  if (sp.flags & PAD_V2_FLAG) tag_size += (size_t)sp.padding;
  if (tag_size > 0x0fffffff) return MP3B200_ERR_BUFFER;
  if ((size_t)(cap < 0 ? 0 : cap) < tag_size) return (int)tag_size;      /* like the reference: the size it needs */
  if (!buf) return 0;
  uint8_t* p = buf;
  *p++ = 'I'; *p++ = 'D'; *p++ = '3'; *p++ = 3; *p++ = 0; *p++ = 0;
  const unsigned long body = (unsigned long)tag_size - 10;               /* 28 bits, 7 per byte */
  *p++ = (uint8_t)((body >> 21) & 0x7f); *p++ = (uint8_t)((body >> 14) & 0x7f); *p++ = (uint8_t)((body >> 7) & 0x7f); *p++ = (uint8_t)(body & 0x7f);
  for (const auto& f : sp.frames) {
    memcpy(p, f.id, 4); p += 4;
    p = put32(p, (unsigned long)frame_size(f) - 10);
    *p++ = 0; *p++ = 0;                                                  /* frame flags */
    *p++ = 0;                                                            /* ISO-8859-1 */
    if (f.comment) { *p++ = 'X'; *p++ = 'X'; *p++ = 'X'; *p++ = 0; }     /* language "XXX", empty description */
    memcpy(p, f.text.data(), f.text.size()); p += f.text.size();
  }
  memset(p, 0, buf + tag_size - p);                                      /* padding */
  return (int)tag_size;
}

int mp3b200_id3v1_tag(const mp3b200_id3tag* t, uint8_t* buf, int cap) {
  if (cap < 128) return 128;
  if (!buf) return 0;
  Spec sp;
  const int rc = build_spec(t, &sp);
  if (rc) return rc;
  if (!(sp.flags & CHANGED_FLAG) || (sp.flags & V2_ONLY_FLAG)) return 0;
  const int pad = (sp.flags & SPACE_V1_FLAG) ? ' ' : 0;
  uint8_t* p = buf;
  *p++ = 'T'; *p++ = 'A'; *p++ = 'G';
  p = text_field(p, &sp.title, 30, pad);
  p = text_field(p, &sp.artist, 30, pad);
  p = text_field(p, &sp.album, 30, pad);
  char y[16];
  snprintf(y, sizeof y, "%d", sp.year);
  const std::string year = y;
  p = text_field(p, sp.year != 0 ? &year : nullptr, 4, pad);
  p = text_field(p, &sp.comment, sp.track != 0 ? 28 : 30, pad);
  if (sp.track != 0) { *p++ = 0; *p++ = (uint8_t)sp.track; }            /* version 1.1 */
  *p++ = (uint8_t)sp.genre;
  return 128;
}

const char* mp3b200_id3_genre_name(int index) { return (index >= 0 && index < 148) ? kGenres[index] : nullptr; }

}  // extern "C"
